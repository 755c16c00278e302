// Shared device/host definitions of the SAC update "program": a static list of tile tasks grouped into
// dependency stages.  The same task code runs (a) as one kernel per stage inside a CUDA graph and
// (b) inside ONE persistent cooperative launch with grid barriers between stages.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sacb {

// ---- reference constants ------------------------------------------------------------------
constexpr float kLogStdMin = -20.f, kLogStdMax = 2.f;     // networks_model1.py:74, networks_model2.py:95
constexpr float kSquashEps = 1e-6f;                        // networks_model1.py:96
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;     // torch Normal.log_prob
constexpr float kBeta1 = 0.9f, kBeta2 = 0.999f, kAdamEps = 1e-8f;  // torch.optim.Adam defaults (sac_imp.py:39-41)

// ---- a pointer that is valid for every agent of a population: region base + offset ----------------------
// region 0 = arena (params, targets, Adam state, scalars), region 1 = workspace (minibatch, activations)
struct Ref {
    int64_t v;   // -1 = null ; bit 62 = region ; low bits = float offset
};
__host__ __device__ inline Ref make_ref(int region, int64_t off) { Ref r; r.v = ((int64_t)region << 62) | off; return r; }
__host__ __device__ inline Ref null_ref() { Ref r; r.v = -1; return r; }
__host__ __device__ inline bool is_null(Ref r) { return r.v < 0; }

struct AgentBases {
    float *arena;      // base of agent 0
    float *ws;
    int64_t arena_stride, ws_stride;   // floats between consecutive agents
};
__device__ __forceinline__ float *resolve(Ref r, const AgentBases &b, int agent) {
    if (r.v < 0) return nullptr;
    const int64_t off = r.v & ((1ll << 62) - 1);
    return ((r.v >> 62) & 1) ? b.ws + agent * b.ws_stride + off : b.arena + agent * b.arena_stride + off;
}

// ---- scalars kept per agent in the arena (offsets in floats from the scalar block) ----------------------
enum ScalarSlot {
    SC_LOG_ALPHA = 0, SC_LOG_ALPHA_M, SC_LOG_ALPHA_V,
    SC_ALPHA0, SC_ALPHA1,            // alpha used by step t = SC_ALPHA0 + (n_updates & 1); written for t+1 into the other
    SC_STEP_POLICY, SC_STEP_Q1, SC_STEP_Q2, SC_STEP_ALPHA, SC_N_UPDATES,   // stored as int32 bit patterns
    SC_LOSS_Q1, SC_LOSS_Q2, SC_LOSS_PI, SC_LOSS_ALPHA,
    SC_COUNT = 16
};

// ---- GEMM operand ---------------------------------------------------------------------------------------
// logical operand X[R, K] (R = M for A, N for B).  mn_major = 0: stored row-major [R, K] (K contiguous);
// mn_major = 1: stored row-major [K, R] (R contiguous) -> the smem fill transposes.
// xform = 1: value = (src > 0) ? rvec[storage_row] * cvec[storage_col] : 0   (implicit dL/dh of the last
// hidden layer: dq[b] * w_out[n] * relu'(h[b,n]); never materialised)
struct Operand {
    Ref ptr;
    int32_t ld;
    int32_t mn_major;
    int32_t xform;
    int32_t pad;
    Ref rvec, cvec;
};

enum TaskType : int32_t {
    T_GEMM = 0,
    T_GATHER,        // replay ring rows -> minibatch matrices
    T_SAMPLE,        // policy head outputs -> tanh-Gaussian sample + log-prob     (networks_model1.py:78-99)
    T_TARGET_LOSS,   // Bellman target + twin critic MSE + dL/dq                   (sac_imp.py:92-105)
    T_ACTOR_LOSS,    // policy loss, min-Q routing, alpha loss + alpha Adam        (sac_imp.py:117-135)
    T_SAMPLE_BWD,    // dL/da -> dL/dmean, dL/dlog_std                             
    T_OUT_ADAM,      // Q output layer: dW=dq^T h, db=sum dq, Adam, Polyak
    T_BIAS_ADAM,     // hidden-layer bias: db = colsum(dh), Adam, Polyak
    T_FINISH         // bump step counters
};

enum Epilogue : int32_t {
    EPI_STORE = 0,       // C = acc
    EPI_BIAS,            // C = acc + bias[n]
    EPI_BIAS_RELU,       // C = relu(acc + bias[n])
    EPI_MASK,            // C = acc * (mask[m,n] > 0)
    EPI_ADAM             // acc = dW tile: Adam on W (+ Polyak into Wt, + optional grad export)
};

struct AdamArgs {
    Ref w, m, v;        // same shape as the GEMM output, ld = N
    Ref wt;             // Polyak target (null for the policy)
    Ref gexp;           // gradient export (null unless SACB_EXPORT_GRADS / data-parallel mode)
    int32_t step_slot;  // ScalarSlot of the optimizer step counter (value BEFORE this step's increment)
    int32_t apply;      // 0 = only export the gradient (data-parallel backward), 1 = apply Adam
    float lr, tau;
};

struct Task {
    int32_t type;
    int32_t tile_begin;       // first tile of this task inside its stage
    int32_t n_tiles;
    int32_t tiles_m, tiles_n;
    // --- GEMM
    Operand A, B;
    int32_t M, N, K;
    int32_t epi;
    Ref C; int32_t ldc;
    int32_t accumulate;       // EPI_STORE: C += acc
    Ref bias;
    Ref mask; int32_t ld_mask;
    AdamArgs adam;
    // --- elementwise tasks: generic slots (meaning depends on type, see tasks.cuh)
    Ref p[24];
    int32_t i[8];
    float f[6];
};

struct Stage {
    int32_t task_begin, task_end;
    int32_t n_tiles;          // per agent
    int32_t pad;
};

struct Program {
    const Task *tasks;
    const Stage *stages;
    int32_t n_stages;
    int32_t n_agents;
    AgentBases bases;
    Ref scalars;              // arena ref of the scalar block
    unsigned int *barrier;    // grid barrier counter (persistent mode)
    // replay ring (T_GATHER)
    const float *ring; int64_t ring_agent_stride; int32_t ring_row; int32_t pad0;
    const int32_t *slots;     // physical ring slots of the minibatch rows [n_agents, B]
    int32_t slots_stride; int32_t pad1;
    int32_t *error_flag;      // set by watchdogs (mbarrier / grid barrier timeouts)
    unsigned long long *trace; // optional [grid][8] globaltimer stamps of each CTA's first tile (profiling aid), or null
};

// ---- tile geometry ----------------------------------------------------------------------------------------
constexpr int kThreads = 512;
// FFMA path
constexpr int kSM = 64, kSN = 64, kSK = 16;
// tcgen05 path: 128 x kTN output tile, K blocks of 32 fp32 (=128 B, one SWIZZLE_128B row)
constexpr int kTM = 128, kTN = 64, kTK = 32, kTStages = 4;
constexpr int kTcStageBytes = (kTM + kTN) * kTK * 4;              // 24 KB
constexpr int kTcXformBytes = (128 + 2048) * 4;                   // rank-1 transform vectors (gemm.cuh: xr, xk)
__host__ __device__ constexpr int tc_smem_bytes(int split) { return kTStages * split * kTcStageBytes + kTcXformBytes + 1024; }
constexpr int kTcSmemBytes = tc_smem_bytes(1);
constexpr int kSimtSmemBytes = 2 * kSK * (kSM + 4) * 4;

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace sacb
