// Non-GEMM tile tasks of the SAC update program (sampling, losses, bias / output-layer optimiser steps).
// Slot meaning of Task::p / i / f is documented per task; the host builder (program.cu) fills them.
#pragma once
#include "gemm.cuh"

namespace sacb {

// ---- Philox4x32-10 (Salmon et al. 2011) for production-mode eps draws --------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream, uint32_t step, uint32_t row, uint32_t col) {
    uint32_t c[4] = {row, col, step, stream};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float row_dot(const float *h, const float *w, int n, int lane) {
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s = fmaf(ldcg(h + j), ldcg(w + j), s);
    return warp_sum(s);
}

// tanh-Gaussian sample of one action component: networks_model1.py:83-96 == networks_model2.py:104-117
struct SampleElem { float action, logp, y, std, in_range; };
__device__ __forceinline__ SampleElem sample_elem(float mean, float ls_raw, float eps, float scale, float bias) {
    SampleElem o;
    const float ls = fminf(fmaxf(ls_raw, kLogStdMin), kLogStdMax);     // torch.clamp(log_std, -20, 2)
    o.in_range = (ls_raw >= kLogStdMin && ls_raw <= kLogStdMax) ? 1.f : 0.f;
    o.std = expf(ls);
    const float x = mean + eps * o.std;                                  // Normal.rsample
    o.y = tanhf(x);
    o.action = o.y * scale + bias;
    const float var = o.std * o.std;
    const float d = x - mean;
    float lp = -(d * d) / (2.f * var) - logf(o.std) - kLogSqrt2Pi;      // Normal.log_prob
    lp -= logf(scale * (1.f - o.y * o.y) + kSquashEps);
    o.logp = lp;
    return o;
}

// T_GATHER: p0=Xall [3B,ldx] ; p1=r ; p2=d ; i0=B i1=obs i2=act i3=ldx ; ring row = [s | s2 | a | r | d] (16 B aligned)
//   rows [0,B) <- (s2, .)   rows [B,2B) <- (s, a)   rows [2B,3B) <- (s, .)
// one warp per sampled row: the whole row is fetched with 128-bit streaming loads issued back to back
__device__ __forceinline__ void task_gather(const Task &t, int tile, const Program &P, int agent) {
    const int B = t.i[0], obs = t.i[1], act = t.i[2], ldx = t.i[3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = tile * (kThreads / 32) + warp;
    if (b >= B) return;
    float *X = resolve(t.p[0], P.bases, agent);
    const int slot = P.slots[(int64_t)agent * P.slots_stride + b];
    const float4 *row = reinterpret_cast<const float4 *>(P.ring + agent * P.ring_agent_stride + (int64_t)slot * P.ring_row);
    float *x2 = X + (int64_t)b * ldx, *x1 = X + (int64_t)(B + b) * ldx, *x3 = X + (int64_t)(2 * B + b) * ldx;
    float *rr = resolve(t.p[1], P.bases, agent), *dd = resolve(t.p[2], P.bases, agent);
    const int nvec = P.ring_row >> 2;
    for (int v0 = 0; v0 < nvec; v0 += 32 * 8) {
        float4 buf[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int v = v0 + u * 32 + lane;
            buf[u] = v < nvec ? __ldcs(row + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int v = v0 + u * 32 + lane;
            if (v >= nvec) continue;
            const float e[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = 4 * v + q;
                if (j < obs) { x1[j] = e[q]; x3[j] = e[q]; }
                else if (j < 2 * obs) x2[j - obs] = e[q];
                else if (j < 2 * obs + act) x1[obs + (j - 2 * obs)] = e[q];
                else if (j == 2 * obs + act) rr[b] = e[q];
                else if (j == 2 * obs + act + 1) dd[b] = e[q];
            }
        }
    }
}

// T_SAMPLE: p0=head_raw [2B,2A] (mean | log_std_raw) ; p1=eps [2B,A] (null -> Philox) ; p2=Xall ; p3=logp [2B]
//   i0=B i1=A i2=obs i3=ldx ; f0=scale f1=bias.  row j<B: next-state sample -> X2[j,obs:] ; j>=B: current -> X3[j-B,obs:]
__device__ __forceinline__ void task_sample(const Task &t, int tile, const Program &P, int agent, const float *scalars, uint64_t seed) {
    const int B = t.i[0], A = t.i[1], obs = t.i[2], ldx = t.i[3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = tile * (kThreads / 32) + warp;
    if (j >= 2 * B) return;
    const float *head = resolve(t.p[0], P.bases, agent) + (int64_t)j * 2 * A;
    float *eps = resolve(t.p[1], P.bases, agent);
    float *X = resolve(t.p[2], P.bases, agent);
    float *dst = (j < B) ? X + (int64_t)j * ldx + obs : X + (int64_t)(2 * B + (j - B)) * ldx + obs;
    const uint32_t step = (uint32_t)__float_as_int(ldcg(scalars + SC_N_UPDATES));
    float lp = 0.f;
    for (int a = lane; a < A; a += 32) {
        float e;
        if (t.i[4]) {   // production mode: draw on device and keep the draw for the backward pass
            e = philox_normal(seed, (uint32_t)agent, step, (uint32_t)j, (uint32_t)a);
            eps[(int64_t)j * A + a] = e;
        } else {
            e = ldcg(eps + (int64_t)j * A + a);
        }
        const SampleElem s = sample_elem(ldcg(head + a), ldcg(head + A + a), e, t.f[0], t.f[1]);
        dst[a] = s.action;
        lp += s.logp;
    }
    lp = warp_sum(lp);
    if (lane == 0) resolve(t.p[3], P.bases, agent)[j] = lp;
}

// dot products of one activation row with up to 4 weight vectors, 128-bit loads, all issued before use
template <int NV>
__device__ __forceinline__ void row_dots(const float *(&h)[NV], const float *(&w)[NV], int64_t row_off, int n, int lane, float (&out)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = 0.f;
    for (int j = lane * 4; j < n; j += 128) {
        float4 hv[NV], wv[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            hv[k] = __ldcg(reinterpret_cast<const float4 *>(h[k] + row_off + j));
            wv[k] = __ldcg(reinterpret_cast<const float4 *>(w[k] + j));
        }
#pragma unroll
        for (int k = 0; k < NV; k++) out[k] += hv[k].x * wv[k].x + hv[k].y * wv[k].y + hv[k].z * wv[k].z + hv[k].w * wv[k].w;
    }
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = warp_sum(out[k]);
}

// T_TARGET_LOSS (one warp per row, 8 rows per tile): Bellman target + critic MSE terms + dL/dq   (sac_imp.py:92-105)
//   p0,p1 = last hidden activations of q1_target,q2_target on (s2,a2) [B,H] ; p2,p3 = of q1,q2 on (s,a)
//   p4..p7 = output-layer weights [H] of q1t,q2t,q1,q2 ; p8..p11 = their biases [1]
//   p12=r p13=d p14=logp_next p15=is_weights(null -> 1) ; outputs p16=y p17=dq1 p18=dq2 p19=td (|q1-y|)
//   p[20],p[21] = snapshot copies of the q1,q2 output weights read by the rank-1 operand transforms
//   p[22] = per-tile partial sums of w*(q-y)^2 [n_tiles, 2] (summed in tile order by T_FINISH: deterministic)
//   i0=B i1=H ; f0=gamma.   H % 4 == 0 (checked at create)
__device__ __forceinline__ void task_target_loss(const Task &t, int tile, const Program &P, int agent, const float *scalars, float *smem) {
    const int B = t.i[0], H = t.i[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
    const float *h[4], *w[4];
    for (int k = 0; k < 4; k++) { h[k] = resolve(t.p[k], P.bases, agent); w[k] = resolve(t.p[4 + k], P.bases, agent); }
    if (tile == 0) {
        float *snap1 = resolve(t.p[20], P.bases, agent), *snap2 = resolve(t.p[21], P.bases, agent);
        for (int j = threadIdx.x; j < H; j += kThreads) { snap1[j] = ldcg(w[2] + j); snap2[j] = ldcg(w[3] + j); }
    }
    const int b = tile * nw + warp;
    float l1 = 0.f, l2 = 0.f;
    if (b < B) {
        float q[4];
        row_dots<4>(h, w, (int64_t)b * H, H, lane, q);
        if (lane == 0) {
            for (int k = 0; k < 4; k++) q[k] += ldcg(resolve(t.p[8 + k], P.bases, agent));
            const int n_upd = __float_as_int(ldcg(scalars + SC_N_UPDATES));
            const float alpha = ldcg(scalars + SC_ALPHA0 + (n_upd & 1));
            const float *isw = resolve(t.p[15], P.bases, agent);
            const float qn = fminf(q[0], q[1]);
            const float vt = qn - alpha * ldcg(resolve(t.p[14], P.bases, agent) + b);                                   // sac_imp.py:97
            const float yy = ldcg(resolve(t.p[12], P.bases, agent) + b) + (1.f - ldcg(resolve(t.p[13], P.bases, agent) + b)) * t.f[0] * vt;   // :98
            const float wgt = isw ? ldcg(isw + b) : 1.f;
            const float e1 = q[2] - yy, e2 = q[3] - yy;
            resolve(t.p[16], P.bases, agent)[b] = yy;
            resolve(t.p[17], P.bases, agent)[b] = 2.f * wgt * e1 / (float)B;                                          // d mean((q-y)^2) / dq
            resolve(t.p[18], P.bases, agent)[b] = 2.f * wgt * e2 / (float)B;
            resolve(t.p[19], P.bases, agent)[b] = fabsf(e1);
            l1 = wgt * e1 * e1; l2 = wgt * e2 * e2;
        }
    }
    if (lane == 0) { smem[warp] = l1; smem[nw + warp] = l2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < nw; i++) { s1 += smem[i]; s2 += smem[nw + i]; }
        float *part = resolve(t.p[22], P.bases, agent);
        part[2 * tile] = s1; part[2 * tile + 1] = s2;
    }
    __syncthreads();
}

// T_ACTOR_LOSS (one warp per row): policy-loss terms and min-Q routing (sac_imp.py:117-121)
//   p0,p1 = last hidden activations of q1,q2 on (s, a_new) ; p2,p3 = output weights ; p4,p5 = output biases
//   p6=logp_cur ; outputs p7=dqa1 p8=dqa2 ; p9 = per-tile partials [n_tiles,2]: sum(alpha*logp - minq), sum(logp + target_entropy)
//   i0=B i1=H ; f0=target_entropy
__device__ __forceinline__ void task_actor_loss(const Task &t, int tile, const Program &P, int agent, const float *scalars, float *smem) {
    const int B = t.i[0], H = t.i[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
    const float *h[2] = {resolve(t.p[0], P.bases, agent), resolve(t.p[1], P.bases, agent)};
    const float *w[2] = {resolve(t.p[2], P.bases, agent), resolve(t.p[3], P.bases, agent)};
    const int b = tile * nw + warp;
    float pl = 0.f, ent = 0.f;
    if (b < B) {
        float q[2];
        row_dots<2>(h, w, (int64_t)b * H, H, lane, q);
        if (lane == 0) {
            const float q1 = q[0] + ldcg(resolve(t.p[4], P.bases, agent)), q2 = q[1] + ldcg(resolve(t.p[5], P.bases, agent));
            const int n_upd = __float_as_int(ldcg(scalars + SC_N_UPDATES));
            const float alpha = ldcg(scalars + SC_ALPHA0 + (n_upd & 1));
            const float l = ldcg(resolve(t.p[6], P.bases, agent) + b);
            pl = alpha * l - fminf(q1, q2);                                          // sac_imp.py:119-121
            const float sel = q1 < q2 ? 1.f : (q1 == q2 ? 0.5f : 0.f);               // torch.minimum backward
            resolve(t.p[7], P.bases, agent)[b] = -sel / (float)B;
            resolve(t.p[8], P.bases, agent)[b] = -(1.f - sel) / (float)B;
            ent = l + t.f[0];
        }
    }
    if (lane == 0) { smem[warp] = pl; smem[nw + warp] = ent; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < nw; i++) { s1 += smem[i]; s2 += smem[nw + i]; }
        float *part = resolve(t.p[9], P.bases, agent);
        part[2 * tile] = s1; part[2 * tile + 1] = s2;
    }
    __syncthreads();
}

// T_SAMPLE_BWD: p0=da1 p1=da2 [B,A] ; p2=head_raw rows of the current-state sample ; p3=eps_cur ; p4=g_head [B,ldg]
//   i0=B i1=A i2=ldg (row stride of g_head, 16 B aligned) ; f0=scale f1=bias    (closed form of SURVEY 3.3)
__device__ __forceinline__ void task_sample_bwd(const Task &t, int tile, const Program &P, int agent, const float *scalars) {
    const int B = t.i[0], A = t.i[1];
    const int idx = tile * kThreads + threadIdx.x;
    if (idx >= B * A) return;
    const int b = idx / A, a = idx % A;
    const int n_upd = __float_as_int(ldcg(scalars + SC_N_UPDATES));
    const float alpha = ldcg(scalars + SC_ALPHA0 + (n_upd & 1));
    const float *head = resolve(t.p[2], P.bases, agent) + (int64_t)b * 2 * A;
    const float eps = ldcg(resolve(t.p[3], P.bases, agent) + idx);
    const SampleElem s = sample_elem(ldcg(head + a), ldcg(head + A + a), eps, t.f[0], t.f[1]);
    const float da = ldcg(resolve(t.p[0], P.bases, agent) + idx) + ldcg(resolve(t.p[1], P.bases, agent) + idx);
    const float sc = t.f[0], omy2 = 1.f - s.y * s.y, aB = alpha / (float)B;
    const float g_u = da * sc * omy2 + aB * (2.f * sc * s.y * omy2) / (sc * omy2 + kSquashEps);
    float *g = resolve(t.p[4], P.bases, agent) + (int64_t)b * t.i[2];
    g[a] = g_u;
    g[A + a] = (g_u * s.std * eps - aB) * s.in_range;
}

// column sums over the batch for 32 columns per tile: thread (cg = tid & 31, rg = tid >> 5) adds rows rg, rg+G, ...
// (loads unrolled 8 deep), the 8 row groups are then combined in shared memory in a fixed order
template <class F>
__device__ __forceinline__ float colsum32(int B, float *smem, F value_at) {
    constexpr int G = kThreads / 32;       // row groups
    const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
    float acc = 0.f;
    int b = rg;
    for (; b + 7 * G < B; b += 8 * G) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = value_at(b + G * u);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    for (; b < B; b += G) acc += value_at(b);
    smem[rg * 32 + cg] = acc;
    __syncthreads();
    float tot = 0.f;
    if (rg == 0) for (int g = 0; g < G; g++) tot += smem[g * 32 + cg];
    __syncthreads();
    return tot;     // valid for rg == 0
}

// T_OUT_ADAM: Q output layer (Linear(H,1)).  p0=h_L [B,H] p1=dq [B] ; p2..p6 = w,m,v,wt,gexp [H] ; p7..p11 = same for bias
//   i0=B i1=H i2=step_slot i3=apply ; f0=lr f1=tau.   32 columns per tile; tile 0 also does the bias
__device__ __forceinline__ void task_out_adam(const Task &t, int tile, const Program &P, int agent, const float *scalars, float *smem) {
    const int B = t.i[0], H = t.i[1];
    const int n = tile * 32 + (threadIdx.x & 31);
    const float *h = resolve(t.p[0], P.bases, agent), *dq = resolve(t.p[1], P.bases, agent);
    float ss, bs;
    adam_factors(__float_as_int(ldcg(scalars + t.i[2])), t.f[0], ss, bs);
    const float g = colsum32(B, smem, [&](int b) { return n < H ? ldcg(dq + b) * ldcg(h + (int64_t)b * H + n) : 0.f; });
    if (threadIdx.x < 32 && n < H) {
        float *wt = resolve(t.p[5], P.bases, agent), *ge = resolve(t.p[6], P.bases, agent);
        adam_element(g, resolve(t.p[2], P.bases, agent) + n, resolve(t.p[3], P.bases, agent) + n, resolve(t.p[4], P.bases, agent) + n,
                     wt ? wt + n : nullptr, ge ? ge + n : nullptr, t.i[3], ss, bs, t.f[1]);
    }
    if (tile == 0 && threadIdx.x < 32) {
        float gb = 0.f;
        for (int b = threadIdx.x; b < B; b += 32) gb += ldcg(dq + b);
        gb = warp_sum(gb);
        if (threadIdx.x == 0)
            adam_element(gb, resolve(t.p[7], P.bases, agent), resolve(t.p[8], P.bases, agent), resolve(t.p[9], P.bases, agent),
                         resolve(t.p[10], P.bases, agent), resolve(t.p[11], P.bases, agent), t.i[3], ss, bs, t.f[1]);
    }
}

// T_BIAS_ADAM: db[n] = sum_b dh[b,n], dh described by Task::A (K-major [B,N], optional rank-1 transform).
//   p0..p4 = b,m,v,bt,gexp ; i0=B i1=N i2=step_slot i3=apply ; f0=lr f1=tau.   32 columns per tile
__device__ __forceinline__ void task_bias_adam(const Task &t, int tile, const Program &P, int agent, const float *scalars, float *smem) {
    const int B = t.i[0], N = t.i[1];
    const int n = tile * 32 + (threadIdx.x & 31);
    const OperandR A = resolve_operand(t.A, P.bases, agent);
    const float cv = (A.xform && n < N) ? ldcg(A.cvec + n) : 0.f;
    const float g = colsum32(B, smem, [&](int b) {
        if (n >= N) return 0.f;
        const float v = ldcg(A.p + (int64_t)b * A.ld + n);
        return A.xform ? (v > 0.f ? ldcg(A.rvec + b) * cv : 0.f) : v;
    });
    if (threadIdx.x < 32 && n < N) {
        float ss, bs;
        adam_factors(__float_as_int(ldcg(scalars + t.i[2])), t.f[0], ss, bs);
        float *bt = resolve(t.p[3], P.bases, agent), *ge = resolve(t.p[4], P.bases, agent);
        adam_element(g, resolve(t.p[0], P.bases, agent) + n, resolve(t.p[1], P.bases, agent) + n, resolve(t.p[2], P.bases, agent) + n,
                     bt ? bt + n : nullptr, ge ? ge + n : nullptr, t.i[3], ss, bs, t.f[1]);
    }
}

// T_FINISH (one thread): finalise the loss scalars from the per-tile partials (fixed order), the temperature step
// (sac_imp.py:128-135) and the optimizer step counters (+1 each, sac_imp.py:109,113,125,134).
//   p0 = critic partials [nt,2] (null: skip) ; p1 = actor partials [nt,2] (null: skip) ; p2 = exported log_alpha gradient (or null)
//   i0..i3 = bump step_policy,q1,q2,alpha ; i4 = bump n_updates ; i5 = n_tiles of the partial arrays ; i6 = auto_entropy ; i7 = apply
//   f0 = lr ; f1 = B
__device__ __forceinline__ void task_finish(const Task &t, const Program &P, int agent, float *scalars, float *smem) {
    const int nt = t.i[5];
    const float Bf = t.f[1];
    const float *cp = resolve(t.p[0], P.bases, agent), *ap = resolve(t.p[1], P.bases, agent);
    // partials fetched in parallel, then summed in tile order by one thread (deterministic)
    for (int i = threadIdx.x; i < 2 * nt && i < kThreads / 2; i += blockDim.x) {
        smem[i] = cp ? ldcg(cp + i) : 0.f;
        smem[kThreads / 2 + i] = ap ? ldcg(ap + i) : 0.f;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (cp) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < nt; i++) { s1 += smem[2 * i]; s2 += smem[2 * i + 1]; }
        scalars[SC_LOSS_Q1] = s1 / Bf; scalars[SC_LOSS_Q2] = s2 / Bf;
    }
    if (ap) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < nt; i++) { s1 += smem[kThreads / 2 + 2 * i]; s2 += smem[kThreads / 2 + 2 * i + 1]; }
        scalars[SC_LOSS_PI] = s1 / Bf;
        const int n_upd = __float_as_int(scalars[SC_N_UPDATES]);
        float alpha_next = scalars[SC_ALPHA0 + (n_upd & 1)];
        if (t.i[6]) {
            const float la = scalars[SC_LOG_ALPHA];
            const float g = -(s2 / Bf);                                      // d(-mean(log_alpha*(logp+H_t)))/dlog_alpha
            scalars[SC_LOSS_ALPHA] = la * g;
            float ss, bs;
            adam_factors(__float_as_int(scalars[SC_STEP_ALPHA]), t.f[0], ss, bs);
            adam_element(g, &scalars[SC_LOG_ALPHA], &scalars[SC_LOG_ALPHA_M], &scalars[SC_LOG_ALPHA_V], nullptr, resolve(t.p[2], P.bases, agent), t.i[7], ss, bs, 0.f);
            alpha_next = expf(scalars[SC_LOG_ALPHA]);                        // self.alpha = self.log_alpha.exp()
        }
        if (t.i[7]) scalars[SC_ALPHA0 + ((n_upd + 1) & 1)] = alpha_next;
    }
    const int slots[4] = {SC_STEP_POLICY, SC_STEP_Q1, SC_STEP_Q2, SC_STEP_ALPHA};
    for (int k = 0; k < 4; k++)
        if (t.i[k]) scalars[slots[k]] = __int_as_float(__float_as_int(scalars[slots[k]]) + 1);
    if (t.i[4]) scalars[SC_N_UPDATES] = __int_as_float(__float_as_int(scalars[SC_N_UPDATES]) + 1);
}

}  // namespace sacb
