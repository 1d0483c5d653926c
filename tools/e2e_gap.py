"""Where does the gap between the device-resident and the end-to-end step of the default bench come from?  Times four variants of the
step in alternating order (thermal / power-cap drift shows up as an order effect)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from multipitch_architectures_b200.engine import CnnStreamEngine
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import get_plan, C1_HZ
    from multipitch_architectures_b200.libdl.nn_models import deep_cnn_segm_sigmoid
    from tests import synth as HO
    dev = torch.device('cuda', 0)
    model = deep_cnn_segm_sigmoid(**bench.DRCNN_KW, precision='fp16')
    bench.make_weights(model)
    model = model.to(dev).eval()
    eng = CnnStreamEngine(model, chunk=646)
    plan = get_plan(22050, float(C1_HZ / 2 ** (2 / 72)), 512, 36, 6, 5, 1, str(dev))
    host = [torch.from_numpy(HO.synth_clip(i, seconds=30.0)).pin_memory() for i in range(2)]
    devc = [c.to(dev) for c in host]
    out_host = torch.empty(1292, 72).pin_memory()

    def resident(i):
        return eng.predict_audio(devc[i % 2], plan)[0]

    def h2d_only(i):
        return eng.predict_audio(host[i % 2].to(dev, non_blocking=True), plan)[0]

    def d2h_only(i):
        out_host.copy_(eng.predict_audio(devc[i % 2], plan)[0], non_blocking=True)

    def e2e(i):
        out_host.copy_(eng.predict_audio(host[i % 2].to(dev, non_blocking=True), plan)[0], non_blocking=True)

    def timed(fn, steps=8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            for i in range(steps):
                fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    with torch.no_grad():
        for i in range(4):
            resident(i)
            e2e(i)
    for rnd in range(3):
        for name, fn in (('resident', resident), ('e2e', e2e), ('h2d_only', h2d_only), ('d2h_only', d2h_only), ('resident', resident)):
            print(rnd, name, round(timed(fn), 3), 'ms/step')


if __name__ == '__main__':
    main()
