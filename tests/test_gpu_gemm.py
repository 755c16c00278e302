"""TMA + tcgen05 (bf16 hi/lo pairs, TMEM accumulator) tile against the FFMA tile and a host float64 reference,
all three on the same pair operands, through the C ABI.  Covers both operand majors (K-major and the transposed
MN-major UMMA descriptors), ragged M/N/K (TMA zero fill) and an unaligned operand offset."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 512, 512), (256, 512, 365), (512, 365, 256), (16, 34, 48), (256, 17, 512),
                                   (100, 70, 45), (34, 512, 256), (512, 684, 256), (130, 65, 129)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
def test_tc_tile_matches_ffma(M, N, K, a_mn, b_mn):
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm(0, M, N, K, a_mn, b_mn, 0, ctypes.byref(err)))
    # the tensor-core tile drops only the lo*lo term (2^-18 relative per product) and accumulates in a different order
    assert err.value < 2e-5, err.value


@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("r0", [344, 8])
def test_tc_tile_operand_offset(b_mn, r0):
    """B = rows/columns [r0, r0+N) of a wider matrix (TMA: an offset along the contiguous dimension must be 16 B aligned)."""
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm(0, 256, 17, 512, 0, b_mn, r0, ctypes.byref(err)))
    assert err.value < 2e-5, err.value


@pytest.mark.parametrize("bm,bn", [(64, 64), (128, 32), (64, 32)])
@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (256, 512, 365), (100, 70, 45), (130, 65, 129), (512, 34, 512), (64, 32, 64)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
def test_tc_tile_shapes(M, N, K, a_mn, b_mn, bm, bn):
    """64-row tiles (tcgen05 M = 64: rows live in the lower half of each TMEM subpartition) and 32-column tiles."""
    if bn == 32 and b_mn:
        pytest.skip("32-column tiles need a K-major B operand")
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm_tile(0, M, N, K, a_mn, b_mn, 0, bm, bn, ctypes.byref(err)))
    assert err.value < 2e-5, err.value


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("M,N,K,ctas", [(256, 512, 512, 148), (1024, 512, 365, 3), (512, 365, 256, 5), (100, 70, 45, 1), (130, 129, 129, 2),
                                        (34, 512, 256, 148), (512, 684, 1024, 7), (2048, 256, 64, 4)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
def test_stream_tile_matches_ffma(M, N, K, ctas, a_mn, b_mn, bn):
    """Throughput ("stream") form of a GEMM stage: 128 x 64 / 128 x 128 / 128 x 256 tiles (4 / 3 / 2 ring slots), a few resident CTAs walking MANY tiles each
    (the producer / MMA / epilogue roles run ahead of each other across tiles, two TMEM accumulators), all operand majors,
    ragged M / N / K."""
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm_stream(0, M, N, K, a_mn, b_mn, 0, bn, ctas, ctypes.byref(err)))
    assert err.value < 2e-5, err.value


@pytest.mark.parametrize("b_mn,r0", [(0, 344), (1, 8)])
def test_stream_tile_operand_offset(b_mn, r0):
    import humanoid_walking_with_sac_b200 as hw
    N_ = hw._native
    err = ctypes.c_float()
    N_.check(N_.lib().sacb_selftest_gemm_stream(0, 512, 17, 512, 0, b_mn, r0, 64, 2, ctypes.byref(err)))
    assert err.value < 2e-5, err.value
