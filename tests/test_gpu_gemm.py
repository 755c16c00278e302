"""tcgen05 (tf32, TMEM accumulator) tile against the FFMA tile and a host float64 reference, through the C ABI."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (256, 512, 512), (256, 512, 365), (512, 365, 256), (16, 34, 48), (256, 17, 512), (100, 70, 45)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("x3", [0, 1])
def test_tc_tile_matches_ffma(M, N, K, a_mn, b_mn, x3):
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm(0, M, N, K, a_mn, b_mn | (x3 << 1), ctypes.byref(err)))
    # single pass: two operands rounded to tf32 (2^-11 each).  3xTF32: hi/lo operand pairs, fp32-level agreement
    assert err.value < (2e-5 if x3 else 3e-3), err.value
