import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.refshapes import build_model
from tests.weights import fill_state_dict, synth_patches, synth_targets
from multipitch_architectures_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else 'saunet_tiny'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 5
orig = _lib.call
def traced(n, *a):
    try:
        orig(n, *a)
        torch.cuda.synchronize()
    except Exception as e:
        print('FAILED in', n, [tuple(t.shape) if isinstance(t, torch.Tensor) else t for t in a][:40])
        raise
_lib.call = traced
import multipitch_architectures_b200.ops as ops, multipitch_architectures_b200.training as tr, multipitch_architectures_b200.training_unet as tu
ops.call = traced; tr.call = traced; tu.call = traced
scheme = sys.argv[3] if len(sys.argv) > 3 else 'adversarial'
for prec in (['fp32', 'bf16'] if len(sys.argv) > 4 else ['bf16']):
  m = build_model(name, precision=prec)
  m.load_state_dict(fill_state_dict(m.state_dict(), 41, scheme=scheme))
  m.p_dropout = 0.0
  for mod in m.modules():
      if hasattr(mod, 'p_dropout'):
          mod.p_dropout = 0.0
  m = m.cuda().train()
  x, t = synth_patches(B, 19).cuda(), synth_targets(B, 19).cuda()
  y = m(x)
  if isinstance(y, tuple): y = y[0]
  loss = torch.nn.BCELoss()(y, t)
  loss.backward()
  torch.cuda.synchronize()
  print('ok', loss.item())
