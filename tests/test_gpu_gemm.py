"""tcgen05 (tf32, TMEM accumulator) tile against the FFMA tile and a host float64 reference, through the C ABI."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (256, 512, 512), (256, 512, 365), (512, 365, 256), (16, 34, 48), (256, 17, 512), (100, 70, 45)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
def test_tc_tile_matches_ffma(M, N, K, a_mn, b_mn):
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm(0, M, N, K, a_mn, b_mn, ctypes.byref(err)))
    # two operands rounded to tf32 (2^-11 each): |err| ~ 5e-4 * sqrt(K) * |a||b| relative to max |c| ~ sqrt(K)/12
    assert err.value < 3e-3, err.value
