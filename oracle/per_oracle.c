/* CPU restatement (plain C) of the reference's prioritized replay arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product never does.
 *
 * Parity status: PINNED against the live reference (tests/golden/per_*.npz made by
 * tests/golden/make_golden.py from /root/reference/replay_buffer.py:25-90).
 *
 * The reference delegates the arithmetic to numpy (unpinned; 2.3.5 in the build container).  The numpy
 * algorithms restated here, with the reference call site each one serves:
 *   - float32 add.reduce = pairwise summation (numpy/_core/src/umath/loops_utils.h.src::pairwise_sum_FLOAT)
 *       <- `probs.sum()`                                   replay_buffer.py:61
 *   - RandomState.choice(n, size, p) with replacement = float64 cumsum, /= last, searchsorted(side='right')
 *     over RandomState.random_sample draws (numpy/random/mtrand.pyx::choice)
 *       <- `np.random.choice(len, n, p=probs)`             replay_buffer.py:64
 *   - float32 power for p**alpha and (N*p)**(-beta) is libm/SVML dependent (<= 1 ulp apart between CPUs,
 *     SURVEY H6.3): the p**alpha table is an INPUT here, IS weights are compared at <= 2 ulp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* numpy pairwise_sum_FLOAT, unit stride (PW_BLOCKSIZE = 128, 8 accumulators). */
float per_oracle_pairwise_sum_f32(const float *a, int64_t n)
{
    if (n < 8) {
        float res = 0.f;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        volatile float r[8];
        float res;
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return per_oracle_pairwise_sum_f32(a, n2) + per_oracle_pairwise_sum_f32(a + n2, n - n2);
    }
}

/* beta annealing: replay_buffer.py:53 (frame is the value BEFORE the += 1 at :54). */
double per_oracle_beta(double beta_start, int64_t beta_frames, int64_t frame)
{
    double b = beta_start + (double)frame * (1.0 - beta_start) / (double)beta_frames;
    return b < 1.0 ? b : 1.0;
}

/* replay_buffer.py:57-68 downstream of the p**alpha table.
 *   p_alpha[n]  float32 priorities**alpha
 *   u[k]        float64 uniforms in [0,1) (what RandomState.random_sample(k) would return)
 *   idx_out[k]  int64, weights_out[k] float32 (already divided by the batch max)
 *   probs_out   optional float32[n] (NULL to skip), cdf_out optional float64[n]
 * returns 0, or -1 on allocation failure. */
int per_oracle_sample(const float *p_alpha, int64_t n, const double *u, int64_t k, double beta,
                      int64_t *idx_out, float *weights_out, float *probs_out, double *cdf_out)
{
    float *probs = probs_out ? probs_out : (float *)malloc(sizeof(float) * (size_t)n);
    double *cdf = cdf_out ? cdf_out : (double *)malloc(sizeof(double) * (size_t)n);
    if (!probs || !cdf) return -1;
    const float total = per_oracle_pairwise_sum_f32(p_alpha, n);       /* :61 probs.sum() */
    for (int64_t i = 0; i < n; i++) probs[i] = p_alpha[i] / total;      /* :61 probs /= sum  (float32) */
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) { acc += (double)probs[i]; cdf[i] = acc; }   /* p.cumsum() in float64 */
    const double last = cdf[n - 1];
    for (int64_t i = 0; i < n; i++) cdf[i] /= last;                     /* cdf /= cdf[-1] */
    for (int64_t j = 0; j < k; j++) {                                   /* searchsorted(u, side='right') */
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (u[j] < cdf[mid]) hi = mid; else lo = mid + 1;
        }
        idx_out[j] = lo;
    }
    float wmax = -INFINITY;
    const float nb = -(float)beta;                                      /* python float -> float32 (NEP 50) */
    for (int64_t j = 0; j < k; j++) {
        int64_t i = idx_out[j] < n ? idx_out[j] : n - 1;
        float w = powf((float)n * probs[i], nb);                        /* :67 */
        weights_out[j] = w;
        if (w > wmax) wmax = w;
    }
    for (int64_t j = 0; j < k; j++) weights_out[j] /= wmax;             /* :68 */
    if (!probs_out) free(probs);
    if (!cdf_out) free(cdf);
    return 0;
}

/* replay_buffer.py:84-87: sequential, later duplicates win; float64 add of 1e-6, stored as float32. */
void per_oracle_update_priorities(float *priorities, const int64_t *idx, const float *td, int64_t k)
{
    for (int64_t j = 0; j < k; j++) priorities[idx[j]] = (float)((double)td[j] + 1e-6);
}

/* replay_buffer.py:36-46: priority given to a pushed transition = max over the WHOLE capacity array
 * (zeros for never-written slots) or 1.0 while the buffer is empty; returns the new write position. */
int64_t per_oracle_push(float *priorities, int64_t capacity, int64_t len_before, int64_t pos)
{
    float mx = 1.0f;
    if (len_before > 0) {
        mx = priorities[0];
        for (int64_t i = 1; i < capacity; i++) if (priorities[i] > mx) mx = priorities[i];
    }
    priorities[pos] = mx;
    return (pos + 1) % capacity;
}

/* glibc powf stand-in for `priorities ** alpha` (replay_buffer.py:60); NOT bit-exact with numpy's SVML path. */
void per_oracle_pow_alpha(const float *priorities, int64_t n, double alpha, float *out)
{
    const float a = (float)alpha;
    for (int64_t i = 0; i < n; i++) out[i] = powf(priorities[i], a);
}
