// Micro-benchmark: how fast does ONE SM ingest GEMM operand K blocks from L2 into shared memory on B200?
// (the K loop of every latency-form GEMM stage is paced by this: profiles/r02_summary.md)
//   mode 0: the product's form: 4-D tensor-map boxes {64 bf16, rows, 2 planes, 1}, SWIZZLE_128B, one box per operand per K block
//   mode 1: the same bytes as 1-D bulk copies of contiguous (pre-tiled) memory, one per operand per K block
//   mode 2: boxes {64, rows, 1 plane, 1}: four instructions per K block
//   mode 4: mode 0 with the two tensor maps read from GLOBAL memory (where the product keeps them: inside the task records)
//   mode 3: mode 0 with both operands fetched from ONE box (rows = rowsA + rowsB of the A matrix): one instruction per K block
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tools/tma_rate.cu      run: ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kStages = 8;      // barrier slots; a run uses p.nst = min(8, 192 KB / K-block bytes) of them, like the product's ring
constexpr int kStageBytes = 48 * 1024;      // slot stride; a 128 x 128 K block (64 KB) is run with kStages3 = 3 slots of 64 KB
constexpr int kMaxKb = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Params {
    int mma_m, mma_n, mma_per_kk, mma_only;      // consumer: 0 = plain arrive; else tcgen05.mma of that shape per 16-wide K step (1: A x B, 2: + A x B_lo, 3: + A_lo x B), commit -> empty
    int nst, stage_bytes, dump;
    int mode, nkb, rows_a, rows_b, early;      // early: 0 = prefetch the descriptor 1 us ahead (spin), 1 = issue right away
    const uint8_t *flat;                        // mode 1 source
    const CUtensorMap *gmaps;                   // mode 4: [0] = A, [1] = B in global memory
    unsigned long long *out;                    // [grid][4 + 2 * kMaxKb]
};

__global__ void __launch_bounds__(64, 1) rate_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA1,
                                                     const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmAB, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[kStages], empty[kStages];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    unsigned long long *o = p.out + (size_t)blockIdx.x * (4 + 2 * kMaxKb);
    const unsigned long long t_start = gtime();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (p.mode == 4) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.gmaps[0])) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.gmaps[1])) : "memory");
        } else {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        }
    }
    __syncthreads();
    const int m0 = (blockIdx.x % 8) * 64, n0 = ((blockIdx.x / 8) % 8) * 64;
    const uint32_t bytes = (uint32_t)(p.rows_a + p.rows_b) * 256u;
    if (threadIdx.x == 0) {
        o[0] = t_start;
        if (!p.early) { const unsigned long long t = gtime(); while (gtime() - t < 1000) { } }
        o[1] = gtime();
        for (int kb = 0; kb < p.nkb && !p.mma_only; kb++) {
            const int s = kb % p.nst;
            mbar_wait(&empty[s], ((kb / p.nst) & 1) ^ 1);
            if (kb == 0) o[4 + kb] = gtime();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(bytes) : "memory");
            const uint32_t sa = smem_u32(tiles + s * p.stage_bytes), sb = sa + p.rows_a * 256, bar = smem_u32(&full[s]);
            const int k0 = (kb % 8) * 64;
            if (p.mode == 0) {
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(bar), "r"(k0), "r"(m0), "r"(0), "r"(0) : "memory");
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(bar), "r"(k0), "r"(n0), "r"(0), "r"(0) : "memory");
            } else if (p.mode == 4) {
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&p.gmaps[0])), "r"(bar), "r"(k0), "r"(m0), "r"(0), "r"(0) : "memory");
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&p.gmaps[1])), "r"(bar), "r"(k0), "r"(n0), "r"(0), "r"(0) : "memory");
            } else if (p.mode == 1) {
                const uint8_t *ga = p.flat + ((size_t)(blockIdx.x % 16) * 8 + (kb % 8)) * 65536, *gb = ga + 32768;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sa), "l"(ga), "r"((uint32_t)p.rows_a * 256u), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sb), "l"(gb), "r"((uint32_t)p.rows_b * 256u), "r"(bar) : "memory");
            } else if (p.mode == 2) {
                for (int pl = 0; pl < 2; pl++) {
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(sa + pl * p.rows_a * 128), "l"(reinterpret_cast<uint64_t>(&tmA1)), "r"(bar), "r"(k0), "r"(m0), "r"(pl), "r"(0) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(sb + pl * p.rows_b * 128), "l"(reinterpret_cast<uint64_t>(&tmB1)), "r"(bar), "r"(k0), "r"(n0), "r"(pl), "r"(0) : "memory");
                }
            } else {
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmAB)), "r"(bar), "r"(k0), "r"(m0), "r"(0), "r"(0) : "memory");
            }
        }
        o[2] = gtime();
    } else if (threadIdx.x >= 32) {
        __shared__ uint32_t s_tmem;
        uint32_t tmem = 0;
        if (p.mma_per_kk) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem = s_tmem;
        }
        if (threadIdx.x == 32) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.mma_n >> 3) << 17) | ((uint32_t)(p.mma_m >> 4) << 24);
            auto desc = [](uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); };
            long long c0[8], c1[8], c2[8];
            for (int kb = 0; kb < p.nkb; kb++) {
                const int s = kb % p.nst;
                if (!p.mma_only) mbar_wait(&full[s], (kb / p.nst) & 1);
                if (p.mma_per_kk) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (kb == 0 || kb == p.nkb - 1) o[4 + kMaxKb + kb] = gtime();      // first and last arrival only: a stamp (timer read + global store) costs ~0.1 us
                c0[kb & 7] = clock64();
                if (p.mma_per_kk) {
                    const uint32_t sa = smem_u32(tiles + s * p.stage_bytes), sb = sa + p.rows_a * 256;
                    for (int kk = 0; kk < 4; kk++) {
                        const uint64_t da = desc(sa + kk * 32), da_lo = desc(sa + p.rows_a * 128 + kk * 32), db = desc(sb + kk * 32), db_lo = desc(sb + p.rows_b * 128 + kk * 32);
                        for (int j = 0; j < p.mma_per_kk; j++) {
                            const uint64_t a = j == 2 ? da_lo : da, b = j == 1 ? db_lo : db;
                            const uint32_t acc = (kb | kk | j) ? 1u : 0u;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
                        }
                    }
                    c1[kb & 7] = clock64();
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
                    c2[kb & 7] = clock64();
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
                }
            }
            if (p.mma_per_kk) {      // all MMAs retired: the last commit has arrived
                const int kb = p.nkb - 1, s = kb % p.nst;
                mbar_wait(&empty[s], (kb / p.nst) & 1);
            }
            o[3] = gtime();
            if (p.mma_per_kk && blockIdx.x == 0 && p.dump) {
                const long long c3 = clock64();
                for (int kb = 0; kb < 8; kb++) printf("   kb %d: wait-done +%lld | mma issue %lld | commit %lld clk\n", kb, c0[kb] - c0[0], c1[kb] - c0[kb], c2[kb] - c1[kb]);
                printf("   all retired +%lld clk\n", c3 - c0[0]);
            }
        }
        __syncwarp();
        if (p.mma_per_kk) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}


// ---- tcgen05.mma rate with both operands in shared memory (SS mode), by shape: back-to-back issue from one thread, one commit at the end ----
template <int M, int N, int PER_KK>
__global__ void __launch_bounds__(32, 1) mma_rate_kernel(long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (64 * 1024) / 16; i += 32) reinterpret_cast<uint4 *>(tiles)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    constexpr int kKb = 32;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sa = smem_u32(tiles), sb = sa + 32 * 1024;      // A planes: hi at 0, lo at +16 KB; B planes: hi at 0, lo at +16 KB (N <= 128) or one 256-row operand
        auto desc = [](uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); };
        const long long t0 = clock64();
#pragma unroll 1
        for (int kb = 0; kb < kKb; kb++) {
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const uint64_t da = desc(sa + kk * 32), da_lo = desc(sa + 16384 + kk * 32), db = desc(sb + kk * 32), db_lo = desc(sb + 16384 + kk * 32);
#pragma unroll
                for (int j = 0; j < PER_KK; j++) {
                    const uint64_t a = j == 2 ? da_lo : da, b = j == 1 ? db_lo : db;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(kb | kk | j)) : "memory");
                }
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}
template <int M, int N, int PER_KK>
static void run_mma_rate(const char *name, long long *d) {
    CK(cudaFuncSetAttribute(mma_rate_kernel<M, N, PER_KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 1024));
    long long h[2] = {0, 0}, best = 1ll << 60, best_issue = 0;
    for (int r = 0; r < 5; r++) {
        mma_rate_kernel<M, N, PER_KK><<<1, 32, 65 * 1024>>>(d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        if (h[1] < best) { best = h[1]; best_issue = h[0]; }
    }
    const double n = 32.0 * 4 * PER_KK, bytes = (M + N) * 32.0;      // operand bytes one instruction reads: (M + N) rows x 16 bf16
    printf("%-44s | %6.1f clk per MMA (issue loop alone %6.1f) | %5.0f B of operands -> %5.1f B/clk | %6.0f flop/clk\n", name, best / n, best_issue / n, bytes, bytes / (best / n),
           2.0 * M * N * 16 / (best / n));
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
    const int rows = 512, cols = 512;      // a 512 x 512 pair matrix: two bf16 planes
    uint8_t *pm = nullptr, *flat = nullptr; unsigned long long *d_out = nullptr;
    CK(cudaMalloc(&pm, (size_t)rows * cols * 2 * 2 * 2)); CK(cudaMemset(pm, 0, (size_t)rows * cols * 2 * 2 * 2));
    CK(cudaMalloc(&flat, 16 * 8 * 65536)); CK(cudaMemset(flat, 0, 16 * 8 * 65536));
    CUtensorMap *gmaps = nullptr; CK(cudaMalloc(&gmaps, 2 * sizeof(CUtensorMap)));
    CK(cudaMalloc(&d_out, sizeof(unsigned long long) * 148 * (4 + 2 * kMaxKb)));
    auto make = [&](int box_rows, int planes, CUtensorMapL2promotion promo) {
        CUtensorMap tm;
        const cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 2, 1};
        const cuuint64_t strides[3] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2, (cuuint64_t)rows * cols * 4};
        const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, (cuuint32_t)planes, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, pm, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); exit(1); }
        return tm;
    };
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kStageBytes + 1024));
    const int nkb = 8;
    printf("== operand ingest: TMA K blocks into a 192 KB ring, consumer = plain mbarrier arrive (no MMA)\n");
    printf("rowsA rowsB slots ctas | per-K-block arrival interval us | KB per K block -> B/clk at 1.965 GHz   (8 K blocks; slots = K blocks requested before the first one is consumed)\n");
    struct V { int ra, rb; };
    const V vs[] = {{64, 32}, {64, 64}, {128, 64}, {128, 128}};
    for (const V &v : vs)
    for (int deep = 0; deep < 2; deep++)
    for (int ctas : {1, 8, 128, 148}) {
        const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        CUtensorMap tA = make(v.ra, 2, promo), tB = make(v.rb, 2, promo), tA1 = make(v.ra, 1, promo), tB1 = make(v.rb, 1, promo), tAB = make(64, 2, promo);
        std::vector<unsigned long long> h(148 * (4 + 2 * kMaxKb));
        double iv = 0;
        const int reps = 5;
        const int kb_bytes = (v.ra + v.rb) * 256;
        const int sbytes = deep ? kb_bytes : std::max(kb_bytes, kStageBytes), nst = deep ? std::min(8, 4 * kStageBytes / sbytes) : std::min(4, 4 * kStageBytes / sbytes);
        if (deep && nst <= 4 && sbytes == std::max(kb_bytes, kStageBytes)) continue;      // same configuration as the shallow run
        Params p{0, 0, 0, 0, nst, sbytes, 0, 0, nkb, v.ra, v.rb, 1, flat, gmaps, d_out};
        for (int r = 0; r < reps + 2; r++) {
            rate_kernel<<<ctas, 64, 4 * kStageBytes + 1024>>>(tA, tB, tA1, tB1, tAB, p);
            CK(cudaDeviceSynchronize());
            if (r < 2) continue;
            CK(cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * ctas * (4 + 2 * kMaxKb), cudaMemcpyDeviceToHost));
            for (int c = 0; c < ctas; c++) {
                const unsigned long long *o = &h[(size_t)c * (4 + 2 * kMaxKb)];
                iv += (double)(o[4 + kMaxKb + nkb - 1] - o[4 + kMaxKb]) / (nkb - 1);
            }
        }
        iv = iv / ((double)reps * ctas) * 1e-3;
        printf("%4d %4d %4d %4d | %.3f | %.0f KB -> %.1f B/clk\n", v.ra, v.rb, nst, ctas, iv, kb_bytes / 1024.0, kb_bytes / (iv * 1965.0));
    }
    printf("== tcgen05.mma kind::f16, cta_group::1, both operands in shared memory (K-major, SWIZZLE_128B), 128 K steps back to back\n");
    long long *d_mma = nullptr; CK(cudaMalloc(&d_mma, 16));
    run_mma_rate<64, 64, 1>("M64 N64, same operands", d_mma);
    run_mma_rate<64, 64, 3>("M64 N64 x3 (hi/lo products, 64 x 64 tile)", d_mma);
    run_mma_rate<128, 32, 3>("M128 N32 x3", d_mma);
    run_mma_rate<128, 64, 1>("M128 N64", d_mma);
    run_mma_rate<128, 64, 3>("M128 N64 x3 (128 x 64 tile)", d_mma);
    run_mma_rate<128, 128, 1>("M128 N128", d_mma);
    run_mma_rate<128, 128, 3>("M128 N128 x3 (stream tile)", d_mma);
    run_mma_rate<128, 256, 1>("M128 N256", d_mma);
    return 0;
}
