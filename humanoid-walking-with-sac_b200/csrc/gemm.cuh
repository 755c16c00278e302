// GEMM tile workers of the SAC update program.
//   gemm_tile_ffma : fp32 FFMA tile (strict-parity mode, and the checker of the tensor-core tile)
//   gemm_tile_tc   : tcgen05.mma kind::tf32 tile, accumulator in TMEM, operands staged in shared memory in
//                    the canonical K-major SWIZZLE_128B layout (fp32 rows of 128 B)
// Both compute  C[M,N] = epilogue( A[M,K] * B[N,K]^T )  for one output tile of a Task.
#pragma once
#include "common.cuh"

namespace sacb {

// ---- resolved (per agent) views ---------------------------------------------------------------------------
struct OperandR {
    const float *p;
    int ld, mn_major, xform;
    const float *rvec, *cvec;
};

struct EpiR {
    int epi, M, N;
    float *C; int ldc; int accumulate;
    const float *bias;
    const float *mask; int ld_mask;
    float *w, *m, *v, *wt, *gexp;
    int apply;
    float step_size, bc2_sqrt, tau;
};

__device__ __forceinline__ float ldcg(const float *p) { return __ldcg(p); }

__device__ __forceinline__ OperandR resolve_operand(const Operand &o, const AgentBases &b, int agent) {
    OperandR r;
    r.p = resolve(o.ptr, b, agent);
    r.ld = o.ld; r.mn_major = o.mn_major; r.xform = o.xform;
    r.rvec = resolve(o.rvec, b, agent);
    r.cvec = resolve(o.cvec, b, agent);
    return r;
}

// value of the logical operand at (r, k); caller guarantees in-bounds
__device__ __forceinline__ float operand_at(const OperandR &o, int r, int k) {
    const int srow = o.mn_major ? k : r, scol = o.mn_major ? r : k;
    float v = ldcg(o.p + (int64_t)srow * o.ld + scol);
    if (o.xform) v = v > 0.f ? ldcg(o.rvec + srow) * ldcg(o.cvec + scol) : 0.f;
    return v;
}

// Adam bias corrections for step t (torch/optim/adam.py::_single_tensor_adam): python floats = double
__device__ __forceinline__ void adam_factors(int step_before, float lr, float &step_size, float &bc2_sqrt) {
    const double t = (double)(step_before + 1);
    const double bc1 = 1.0 - pow((double)0.9, t);
    const double bc2 = 1.0 - pow((double)0.999, t);
    step_size = (float)((double)lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
}

// one Adam element update (+ Polyak, sac_imp.py:146-152); returns nothing, all state in global memory
__device__ __forceinline__ void adam_element(float g, float *w, float *m, float *v, float *wt, float *gexp, int apply,
                                             float step_size, float bc2_sqrt, float tau) {
    if (gexp) *gexp = g;
    if (!apply) return;
    float mm = __ldcg(m), vv = __ldcg(v), ww = __ldcg(w);
    mm = mm + (1.0f - kBeta1) * (g - mm);                 // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * kBeta2 + (1.0f - kBeta2) * g * g;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + kAdamEps;
    ww = ww - step_size * (mm / denom);
    *m = mm; *v = vv; *w = ww;
    if (wt) *wt = __ldcg(wt) * (1.0f - tau) + ww * tau;  // target <- target*(1-tau) + param*tau
}

__device__ __forceinline__ void epilogue_element(const EpiR &e, int m, int n, float acc) {
    if (m >= e.M || n >= e.N) return;
    switch (e.epi) {
        case EPI_STORE: {
            float *c = e.C + (int64_t)m * e.ldc + n;
            *c = e.accumulate ? __ldcg(c) + acc : acc;
        } break;
        case EPI_BIAS: e.C[(int64_t)m * e.ldc + n] = acc + ldcg(e.bias + n); break;
        case EPI_BIAS_RELU: e.C[(int64_t)m * e.ldc + n] = fmaxf(acc + ldcg(e.bias + n), 0.f); break;
        case EPI_MASK: e.C[(int64_t)m * e.ldc + n] = ldcg(e.mask + (int64_t)m * e.ld_mask + n) > 0.f ? acc : 0.f; break;
        case EPI_ADAM: {
            const int64_t o = (int64_t)m * e.N + n;
            adam_element(acc, e.w + o, e.m + o, e.v + o, e.wt ? e.wt + o : nullptr, e.gexp ? e.gexp + o : nullptr,
                         e.apply, e.step_size, e.bc2_sqrt, e.tau);
        } break;
    }
}

__device__ __forceinline__ EpiR resolve_epilogue(const Task &t, const AgentBases &b, int agent, const float *scalars) {
    EpiR e;
    e.epi = t.epi; e.M = t.M; e.N = t.N;
    e.C = resolve(t.C, b, agent); e.ldc = t.ldc; e.accumulate = t.accumulate;
    e.bias = resolve(t.bias, b, agent);
    e.mask = resolve(t.mask, b, agent); e.ld_mask = t.ld_mask;
    e.w = e.m = e.v = e.wt = e.gexp = nullptr; e.apply = 0; e.step_size = e.bc2_sqrt = 0.f; e.tau = 0.f;
    if (t.epi == EPI_ADAM) {
        e.w = resolve(t.adam.w, b, agent); e.m = resolve(t.adam.m, b, agent); e.v = resolve(t.adam.v, b, agent);
        e.wt = resolve(t.adam.wt, b, agent); e.gexp = resolve(t.adam.gexp, b, agent);
        e.apply = t.adam.apply; e.tau = t.adam.tau;
        adam_factors(__float_as_int(ldcg(scalars + t.adam.step_slot)), t.adam.lr, e.step_size, e.bc2_sqrt);
    }
    return e;
}

// ============================================================================================================
// fp32 FFMA tile: 64 x 64 x 16, 256 threads, 4x4 outputs per thread
// ============================================================================================================
__device__ __forceinline__ void gemm_tile_ffma(const Task &t, int tile, const AgentBases &bases, int agent,
                                               const float *scalars, float *smem) {
    const OperandR A = resolve_operand(t.A, bases, agent), B = resolve_operand(t.B, bases, agent);
    const EpiR epi = resolve_epilogue(t, bases, agent, scalars);
    const int tm = tile / t.tiles_n, tn = tile % t.tiles_n;
    const int m0 = tm * kSM, n0 = tn * kSN;
    float(*As)[kSM + 4] = reinterpret_cast<float(*)[kSM + 4]>(smem);
    float(*Bs)[kSN + 4] = reinterpret_cast<float(*)[kSN + 4]>(smem + kSK * (kSM + 4));
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < t.K; k0 += kSK) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int idx = tid + e * kThreads;
            int r, k;
            if (!A.mn_major) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
            As[k][r] = (m0 + r < t.M && k0 + k < t.K) ? operand_at(A, m0 + r, k0 + k) : 0.f;
            if (!B.mn_major) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
            Bs[k][r] = (n0 + r < t.N && k0 + k < t.K) ? operand_at(B, n0 + r, k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSK; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) epilogue_element(epi, m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

// ============================================================================================================
// tcgen05 tile
// ============================================================================================================
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bounded wait: a lost arrival must surface as an error, never as a hung GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *error_flag) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    if (error_flag) atomicExch(error_flag, 1);
    return false;
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs (fp32 bit patterns), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between 8-row
// groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (UMMA::InstrDescriptor): c_format F32=1 [4,6) | a_format TF32=2 [7,10) | b_format [10,13)
// | a_major=b_major=K (0) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte chunk c (4 consecutive k) of row r inside one K-major SWIZZLE_128B operand tile
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// round-to-nearest fp32 -> tf32 (the MMA itself would truncate the low 13 mantissa bits: biased)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

// 4 consecutive-k RAW values of logical row r (k = kk .. kk+3) of an operand; zero outside [0,R) x [0,K).
// Pure loads: nothing here consumes the data, so the loads of k-block kb+2 stay in flight while kb is stored
// and multiplied (the rank-1 transform and the tf32 rounding happen at store time).
__device__ __forceinline__ float4 load_chunk(const OperandR &o, bool vec_ok, int r, int R, int kk, int K) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= R || kk >= K) return v;
    if (!o.mn_major) {
        const float *src = o.p + (int64_t)r * o.ld + kk;
        if (vec_ok && kk + 3 < K) {
            v = __ldcg(reinterpret_cast<const float4 *>(src));
        } else {
            v.x = ldcg(src);
            if (kk + 1 < K) v.y = ldcg(src + 1);
            if (kk + 2 < K) v.z = ldcg(src + 2);
            if (kk + 3 < K) v.w = ldcg(src + 3);
        }
    } else {
        const float *src = o.p + (int64_t)kk * o.ld + r;       // storage (row, col) = (kk+j, r): coalesced over r
        v.x = ldcg(src);
        if (kk + 1 < K) v.y = ldcg(src + o.ld);
        if (kk + 2 < K) v.z = ldcg(src + 2 * (int64_t)o.ld);
        if (kk + 3 < K) v.w = ldcg(src + 3 * (int64_t)o.ld);
    }
    return v;
}

constexpr int kXkMax = 2048;     // longest K a rank-1 transformed operand may have (hidden_dim / batch)

// per-CTA state that survives across tiles (persistent kernel): pipeline position and TMEM base
struct TcState {
    uint32_t tmem_base;
    uint32_t g;            // k-blocks issued so far (stage = g % kTStages)
    uint32_t accum_uses;   // completed tiles (parity of the accumulator barrier)
    uint8_t *tiles;        // 1024-aligned operand ring: kTStages x kSplit x (A 16 KB | B 8 KB)
    float *xr, *xk;        // rank-1 transform vectors of the current tile: by tile row [kTM], by k [kXkMax]
    uint64_t *empty_bar;   // [kTStages]
    uint64_t *accum_bar;
};

constexpr int kAChunks = kTM * (kTK / 4) / kThreads;   // 4 x 16 B per thread per k-block
constexpr int kBChunks = kTN * (kTK / 4) / kThreads;   // 2

struct Regs {
    float4 a[kAChunks];
    float4 b[kBChunks];
};

__device__ __forceinline__ void chunk_coords(bool mn_major, int idx, int rows, int &r, int &c) {
    if (!mn_major) { r = idx >> 3; c = idx & 7; } else { c = idx / rows; r = idx % rows; }
}

__device__ __forceinline__ void load_kblock(Regs &rg, const OperandR &A, const OperandR &B, bool avec, bool bvec,
                                            int m0, int n0, int M, int N, int K, int k0) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int e = 0; e < kAChunks; e++) {
        int r, c; chunk_coords(A.mn_major, tid + e * kThreads, kTM, r, c);
        rg.a[e] = load_chunk(A, avec, m0 + r, M, k0 + 4 * c, K);
    }
#pragma unroll
    for (int e = 0; e < kBChunks; e++) {
        int r, c; chunk_coords(B.mn_major, tid + e * kThreads, kTN, r, c);
        rg.b[e] = load_chunk(B, bvec, n0 + r, N, k0 + 4 * c, K);
    }
}

// hi = rna_tf32(x) ; lo = rna_tf32(x - hi): x = hi + lo up to 2^-22 |x|  (error-compensated "3xTF32")
template <int kSplit>
__device__ __forceinline__ void store_chunk(uint8_t *tile, int tile_bytes, uint32_t off, float4 v) {
    const float4 hi = to_tf32(v);
    *reinterpret_cast<float4 *>(tile + off) = hi;
    if (kSplit == 2) {
        const float4 lo = to_tf32(make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
        *reinterpret_cast<float4 *>(tile + tile_bytes + off) = lo;
    }
}

// stage layout: [A_hi 16K | B_hi 8K] (+ [A_lo | B_lo] when kSplit == 2)
template <int kSplit>
__device__ __forceinline__ void store_kblock(const Regs &rg, const OperandR &A, const OperandR &B, uint8_t *stage, const TcState &st, int k0) {
    const int tid = threadIdx.x;
    constexpr int kHalf = (kTM + kTN) * kTK * 4;
#pragma unroll
    for (int e = 0; e < kAChunks; e++) {
        int r, c; chunk_coords(A.mn_major, tid + e * kThreads, kTM, r, c);
        float4 v = rg.a[e];
        if (A.xform) {      // dq[b] * w_out[n] * relu'(h[b,n]) with the two vectors staged in shared memory
            const float xr = st.xr[r];
            const float *xk = st.xk + k0 + 4 * c;
            v.x = v.x > 0.f ? xr * xk[0] : 0.f; v.y = v.y > 0.f ? xr * xk[1] : 0.f;
            v.z = v.z > 0.f ? xr * xk[2] : 0.f; v.w = v.w > 0.f ? xr * xk[3] : 0.f;
        }
        store_chunk<kSplit>(stage, kHalf, sw128_chunk_off(r, c), v);
    }
#pragma unroll
    for (int e = 0; e < kBChunks; e++) {
        int r, c; chunk_coords(B.mn_major, tid + e * kThreads, kTN, r, c);
        store_chunk<kSplit>(stage, kHalf, kTM * kTK * 4 + sw128_chunk_off(r, c), rg.b[e]);
    }
}

}  // namespace tc

template <int kSplit>
__device__ __forceinline__ void gemm_tile_tc(const Task &t, int tile, const AgentBases &bases, int agent,
                                             const float *scalars, tc::TcState &st, int *error_flag) {
    using namespace tc;
    const OperandR A = resolve_operand(t.A, bases, agent), B = resolve_operand(t.B, bases, agent);
    const EpiR epi = resolve_epilogue(t, bases, agent, scalars);
    const int tm = tile / t.tiles_n, tn = tile % t.tiles_n;
    const int m0 = tm * kTM, n0 = tn * kTN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool avec = !A.mn_major && (A.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.p) & 15) == 0);
    const bool bvec = !B.mn_major && (B.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(B.p) & 15) == 0);
    const int nkb = cdiv(t.K, kTK);
    constexpr uint32_t idesc = make_idesc(kTM, kTN);
    constexpr int kStageBytes = kSplit * kTcStageBytes;
    constexpr int kHalf = kTcStageBytes;

    Regs r0, r1;
    load_kblock(r0, A, B, avec, bvec, m0, n0, t.M, t.N, t.K, 0);
    if (nkb > 1) load_kblock(r1, A, B, avec, bvec, m0, n0, t.M, t.N, t.K, kTK);
    if (A.xform) {   // vectors of the rank-1 operand: one indexed by the tile row, one by k (storage row/col depend on the major)
        const float *by_row = A.mn_major ? A.cvec : A.rvec, *by_k = A.mn_major ? A.rvec : A.cvec;
        for (int i = tid; i < kTM; i += kThreads) st.xr[i] = (m0 + i < t.M) ? ldcg(by_row + m0 + i) : 0.f;
        for (int i = tid; i < nkb * kTK; i += kThreads) st.xk[i] = (i < t.K) ? ldcg(by_k + i) : 0.f;
        __syncthreads();
    }

    auto step = [&](Regs &rg, int kb) {
        const uint32_t g = st.g + kb;
        const uint32_t s = g % kTStages;
        uint8_t *stage = st.tiles + s * kStageBytes;
        if (g >= kTStages) mbar_wait(&st.empty_bar[s], ((g / kTStages) - 1) & 1, error_flag);   // MMAs that read this slot are done
        store_kblock<kSplit>(rg, A, B, stage, st, kb * kTK);
        if (kb + 2 < nkb) load_kblock(rg, A, B, avec, bvec, m0, n0, t.M, t.N, t.K, (kb + 2) * kTK);
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sa = smem_u32(stage), sb = sa + kTM * kTK * 4;
#pragma unroll
            for (int kk = 0; kk < kTK / 8; kk++) {  // UMMA_K = 8 tf32 = 32 B: advance the start address inside the swizzled row
                const uint64_t da = make_desc(sa + kk * 32), db = make_desc(sb + kk * 32);
                if (kSplit == 2) {   // small terms first: a_lo*b_hi + a_hi*b_lo, then a_hi*b_hi
                    umma_tf32(st.tmem_base, make_desc(sa + kHalf + kk * 32), db, idesc, (kb | kk) ? 1u : 0u);
                    umma_tf32(st.tmem_base, da, make_desc(sb + kHalf + kk * 32), idesc, 1u);
                    umma_tf32(st.tmem_base, da, db, idesc, 1u);
                } else {
                    umma_tf32(st.tmem_base, da, db, idesc, (kb | kk) ? 1u : 0u);
                }
            }
            umma_commit(&st.empty_bar[s]);
            if (kb == nkb - 1) umma_commit(st.accum_bar);
        }
    };
    for (int kb = 0; kb < nkb; kb += 2) {
        step(r0, kb);
        if (kb + 1 < nkb) step(r1, kb + 1);
    }
    st.g += nkb;

    // ---- epilogue: TMEM -> registers -> global.  warp w owns TMEM lanes 32*(w%4).., column half w/4
    mbar_wait(st.accum_bar, st.accum_uses & 1, error_flag);
    st.accum_uses++;
    tc_fence_after();
    const int row = (warp & 3) * 32 + lane;
    const int colh = (warp >> 2) * (kTN / 2);
#pragma unroll
    for (int cc = 0; cc < kTN / 2; cc += 16) {
        float v[16];
        tmem_ld16(st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(colh + cc), v);
#pragma unroll
        for (int j = 0; j < 16; j++) epilogue_element(epi, m0 + row, n0 + colh + cc + j, v[j]);
    }
    tc_fence_before();
    __syncthreads();     // all TMEM reads retired before the next tile's first MMA overwrites the accumulator
}

}  // namespace sacb
