"""GPU: the C ABI reports bad arguments as negative return codes with a message (rald_last_error) instead of
crashing or silently falling back — for the entry points added for the eval loop, the fused cross-attention and the
long-context attention."""
import numpy as np
import pytest
import torch

from rald_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _err(name, *args):
    with pytest.raises(_lib.RaldError) as e:
        _lib.call(name, *args)
    return str(e.value)


def test_argument_validation_messages():
    st = _lib.cur_stream()
    x = torch.zeros(1024, device=DEV)
    xi = torch.zeros(16, device=DEV, dtype=torch.int32)
    xd = torch.zeros(16, device=DEV, dtype=torch.float64)
    so = np.ones(6, np.float32)
    v3 = np.ones(3, np.float64)
    r6 = np.ones(6, np.float64)
    # refine: draws must come all together
    msg = _err("rald_refine_queries", x.data_ptr(), xi.data_ptr(), 8, 8, xi.data_ptr(), 0, 0, 0, 2, v3.ctypes.data,
               r6.ctypes.data, so.ctypes.data, x.data_ptr(), st)
    assert "all together" in msg
    assert "aug_num" in _err("rald_refine_queries", x.data_ptr(), xi.data_ptr(), 8, 0, 0, 0, 0, 0, 2, v3.ctypes.data,
                             r6.ctypes.data, so.ctypes.data, x.data_ptr(), st)
    # chamfer: sizes
    assert "bad sizes" in _err("rald_chamfer", x.data_ptr(), xi.data_ptr(), 0, x.data_ptr(), 0, 4, 4, 1, xd.data_ptr(),
                               xd.data_ptr(), st)
    # radar cube prep: channel count / output channels
    assert "bad shape" in _err("rald_radar_cube_prep", x.data_ptr(), 1, 4, 2, 2, 1, 4, 4, 2, 1, 45.0, 1, 2.5,
                               x.data_ptr(), st)
    assert "bad shape" in _err("rald_radar_cube_prep", x.data_ptr(), 1, 4, 2, 2, 3, 4, 4, 3, 1, 45.0, 1, 2.5,
                               x.data_ptr(), st)
    # fused cross-attention: frame window outside the folded operands, rows per frame not a multiple of 128
    assert "frame0" in _err("rald_xattn_fused", x.data_ptr(), x.data_ptr(), x.data_ptr(), 0, x.data_ptr(), 2, 512, 3, 4,
                            st)
    assert "rows/frame" in _err("rald_xattn_fused", x.data_ptr(), x.data_ptr(), x.data_ptr(), 0, x.data_ptr(), 1, 500, 0,
                                1, st)
    assert "null" in _err("rald_xattn_fold", 0, x.data_ptr(), x.data_ptr(), 1, 1, x.data_ptr(), x.data_ptr(), st)
    # long-context attention: context length / scratch
    assert "multiple of 512" in _err("rald_attn_d64_long", x.data_ptr(), 512, x.data_ptr(), 512, x.data_ptr(), 512,
                                     x.data_ptr(), 512, 1, 8, 128, 768, 0.125, x.data_ptr(), x.data_ptr(), st)
    assert "scratch" in _err("rald_attn_d64_long", x.data_ptr(), 512, x.data_ptr(), 512, x.data_ptr(), 512,
                             x.data_ptr(), 512, 1, 8, 128, 1024, 0.125, 0, 0, st)
    torch.cuda.synchronize()      # nothing was launched by the rejected calls; the context is still healthy
    assert float((x + 1).sum()) == 1024.0
