"""`-m "not gpu"`: the C-ABI library builds, loads, and exports every symbol include/mpa.h declares.  No compute."""
import ctypes
import os

import pytest

from multipitch_architectures_b200 import _lib


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from multipitch_architectures_b200.build import build
        build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in include/mpa.h but not exported by libmpa.so: {missing}'


def test_version_and_pure_host_entry_points(lib):
    assert lib.mpa_version() >= 100
    assert lib.mpa_conv_tc_packed_bytes(40, 40, 15, 15, 0) == 17 * 38 * 4096
    assert lib.mpa_conv_tc_packed_bytes(6, 40, 15, 15, 0) == 17 * 8 * 4096
    assert lib.mpa_conv_tc_packed_bytes(40, 200, 15, 15, 0) == 0
    assert lib.mpa_conv_tc_packed_bytes(40, 30, 75, 1, 1) == 75 * 3 * 4096       # conv3 as a J=1 'same' 75x1 convolution
    assert lib.mpa_conv_tc_packed_bytes(40, 80, 3, 3, 2) == 0                    # J*Cout > 128


def test_weight_packing_layout():
    """Host-side packer: tile [k-slice][128 rows][8]; row j*Cout+co holds the filter shifted down by j rows."""
    import numpy as np
    import torch
    from multipitch_architectures_b200 import ops
    rng = np.random.default_rng(0)
    Cin, Cout, KH, KW = 24, 40, 3, 3
    w = torch.from_numpy(rng.standard_normal((Cout, Cin, KH, KW)).astype(np.float32))
    packed = ops.conv_tc_pack(w, 'cpu', ops.FMT_BF16).numpy().view(np.uint16).reshape(KH + 2, -1, 2, 128, 8)
    wb = w.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    mpr = (3 // 2) * KW + (KW + 1) // 2
    assert packed.shape[1] == mpr
    # paired chunks (0,1): mma q = df, k-slice kc -> chunk kc
    for r in range(KH + 2):
        for j in range(3):
            kh = r - j
            for df in range(KW):
                for kc in range(2):
                    got = packed[r, df, kc, j * Cout:(j + 1) * Cout, :]
                    exp = wb[:, kc * 8:(kc + 1) * 8, kh, df] if 0 <= kh < KH else np.zeros((Cout, 8), np.uint16)
                    assert np.array_equal(got, exp)
            # odd chunk 2: taps paired
            for dp in range(2):
                for kc in range(2):
                    df = 2 * dp + kc
                    got = packed[r, KW + dp, kc, j * Cout:(j + 1) * Cout, :]
                    exp = wb[:, 16:24, kh, df] if (0 <= kh < KH and df < KW) else np.zeros((Cout, 8), np.uint16)
                    assert np.array_equal(got, exp)
    assert not packed[:, :, :, 120:, :].any()
    ph = ops.conv_tc_pack(w, 'cpu', ops.FMT_F16).numpy().view(np.uint16).reshape(KH + 2, -1, 2, 128, 8)
    wh = w.half().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(ph[1, 0, 1, 40:80, :], wh[:, 8:16, 0, 0])          # row r=1, j=1 -> kh=0, tap df=0, chunk 1
    tiny = torch.tensor([1e-6, -3e-7, 65504.0, 1e5, 0.0, 6.1e-5, 5.97e-8, 2.0 ** -25]).reshape(8, 1, 1, 1)
    pt = ops.conv_tc_pack(tiny.expand(8, 1, 1, 1).contiguous(), 'cpu', ops.FMT_F16).numpy().view(np.uint16).reshape(-1, 2, 128, 8)
    exp = tiny.reshape(8).half().view(torch.int16).numpy().view(np.uint16)          # host RNE incl. subnormals / overflow
    assert np.array_equal(pt[0, 0, 0:8, 0], exp)


def test_no_cpu_fallback():
    import torch
    from multipitch_architectures_b200 import ops
    with pytest.raises(_lib.MpaError):
        ops.layernorm_cf(torch.zeros(1, 6, 75, 216), torch.ones(6, 216), torch.zeros(6, 216))
