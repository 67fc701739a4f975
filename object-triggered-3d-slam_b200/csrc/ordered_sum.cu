// ordered_sum.cu -- sequential-order FP64 accumulation, bit-identical to the scalar loop, in parallel.
//
// The reference sums triangle areas (TriangleMesh::SamplePointsUniformly) and mean
// neighbour distances (RemoveStatisticalOutliers, A.8) with plain left-to-right loops
//     s = x[0];  for i in 1..n-1:  s = fl(s + x[i])
// and the per-triangle sample counts / the outlier threshold depend on every rounding of that chain.
// A single dependent DADD chain costs ~14 cycles per element (11.7 ms per pass over 1.5 M triangles,
// longer than integrating 500 frames).  This file evaluates the SAME chain exactly, in parallel:
//
//   * all terms are >= 0, so s is non-decreasing and stays inside one binade [2^e, 2^(e+1)) for long
//     runs.  Inside a binade s is an integer multiple m*u of u = 2^(e-52), and as long as the sum stays
//     in the binade   fl(m*u + x) = (m + rn(x/u))*u   where rn rounds to the nearest integer -- the
//     only data dependence on m is the parity rule for exact ties (frac(x/u) == 1/2).  A run of adds
//     is therefore ONE exact integer addition  m += sum_j rn(x_j/u)  unless a tie or a binade crossing
//     occurs in it.
//   * chunks of 256 elements: (A) approximate chunk sums -> (A2) approximate prefix, which predicts
//     each chunk's binade e_c;  (B) per chunk, in parallel, K_c = sum rn(x_j / u_{e_c}) as int64 plus a
//     "complex" flag (tie, term too large, chunk 0, prediction too close to a binade edge);
//     (C) ONE warp walks the chunks in order with the exact accumulator: a simple chunk whose predicted
//     binade matches the exact s and whose m + K_c stays <= 2^53 is a single integer add; any other
//     chunk is replayed with the scalar DADD loop (about 40 of ~6000 chunks for 1.5 M terms: the
//     binade crossings and the occasional tie).  Correctness never rests on the prediction: every
//     shortcut is verified against the exact accumulator, a failed check only costs a scalar replay.
//     (D) prefix mode: simple chunks rebuild their per-element prefixes from the exact carry-in with an
//     integer scan.
// otslam_selftest_ordered_sum compares it with the scalar-loop kernel on adversarial inputs.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace otslam {

constexpr int kOrdStage = 1024;
constexpr int kOsL = 256;                  // elements per chunk: one warp, 8 per lane
constexpr int kOsComplex = INT32_MIN;
constexpr long long kTwo52 = 1ll << 52, kTwo53 = 1ll << 53;

// term of the chain.  mode 0: x   1: x / *div (in-place inclusive prefix)   2: x > 0 ? x : 0   3: x > 0 ? (x - mean)^2 : 0
__device__ __forceinline__ double os_term(int mode, double x, double dv, double mean) {
    if (mode == 1) return __ddiv_rn(x, dv);
    if (mode == 2) return x > 0.0 ? x : 0.0;
    if (mode == 3) { const double d = __dsub_rn(x, mean); return x > 0.0 ? __dmul_rn(d, d) : 0.0; }
    return x;
}

// ---- the scalar chain (small inputs, complex chunks' model, self-test reference): one warp stages
// 1024 terms at a time in SMEM with coalesced loads, lane 0 adds them in index order.
__global__ void __launch_bounds__(32) ordered_accumulate_kernel(double* __restrict__ io, int64_t n, int mode, const double* __restrict__ div,
                                                                double mean, double* __restrict__ out) {
    __shared__ double buf[kOrdStage];
    const int lane = threadIdx.x;
    const double dv = (mode == 1) ? *div : 1.0;
    double nxt[kOrdStage / 32];
    auto load_stage = [&](int64_t base) {
#pragma unroll
        for (int k = 0; k < kOrdStage / 32; ++k) {
            const int64_t i = base + k * 32 + lane;
            nxt[k] = (i < n) ? io[i] : 0.0;
        }
    };
    double acc = 0.0;
    bool first = true;
    load_stage(0);
    for (int64_t base = 0; base < n; base += kOrdStage) {
#pragma unroll
        for (int k = 0; k < kOrdStage / 32; ++k) buf[k * 32 + lane] = os_term(mode, nxt[k], dv, mean);
        if (base + kOrdStage < n) load_stage(base + kOrdStage);     // in flight during the serial walk
        __syncwarp();
        const int cnt = (int)min((int64_t)kOrdStage, n - base);
        if (lane == 0) {
            int j = 0;
            if (first) { acc = buf[0]; j = 1; first = false; }
#pragma unroll 8
            for (; j < cnt; ++j) {
                acc = __dadd_rn(acc, buf[j]);
                if (mode == 1) buf[j] = acc;
            }
        }
        __syncwarp();
        if (mode == 1) {
#pragma unroll
            for (int k = 0; k < kOrdStage / 32; ++k) {
                const int64_t i = base + k * 32 + lane;
                if (i < n) io[i] = buf[k * 32 + lane];
            }
        }
        __syncwarp();
    }
    if (lane == 0 && out) out[0] = acc;
}

// ---- (A) approximate chunk sums; NaN marks a chunk holding a term outside [0, inf)
__global__ void __launch_bounds__(256) os_chunk_sum_kernel(const double* __restrict__ x, int64_t n, int mode, const double* __restrict__ div,
                                                           double mean, double* __restrict__ csum, int64_t n_chunks) {
    const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const int lane = threadIdx.x & 31;
    const double dv = (mode == 1) ? *div : 1.0;
    double acc = 0.0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < kOsL / 32; ++k) {
        const int64_t i = c * kOsL + k * 32 + lane;
        if (i < n) {
            const double v = os_term(mode, x[i], dv, mean);
            bad |= !(v >= 0.0 && v < INFINITY);
            acc += v;
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) csum[c] = bad ? NAN : acc;
}

// ---- (A2) exclusive prefix of the approximate chunk sums (single CTA, tiles of 1024 with a running carry)
__global__ void __launch_bounds__(1024) os_prefix_kernel(const double* __restrict__ csum, double* __restrict__ cpre, int64_t n_chunks) {
    __shared__ double wsum[32];
    __shared__ double carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < n_chunks; base += 1024) {
        const int64_t i = base + t;
        const double v = i < n_chunks ? csum[i] : 0.0;
        double inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            double s = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += u;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const double before = carry_s + (w ? wsum[w - 1] : 0.0) + (inc - v);
        if (i < n_chunks) cpre[i] = before;
        __syncthreads();
        if (t == 1023) carry_s = before + v;
        __syncthreads();
    }
}

// rn(v / 2^(e-52)) for v >= 0 as an integer; flags a tie or a term that does not fit below 2^(e+1)
__device__ __forceinline__ long long os_quantise(double v, int e, bool& irregular) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const int eb = (int)(b >> 52);                       // sign bit is 0 (v >= 0 checked by the caller)
    if (eb == 0) return 0;                               // zero / denormal: far below u/2 (e >= -900)
    const int ev = eb - 1023;
    const unsigned long long mv = (b & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    const int shift = e - ev;
    if (shift < 0) { irregular = true; return 0; }
    if (shift == 0) return (long long)mv;
    if (shift >= 54) return 0;
    const unsigned long long half = 1ull << (shift - 1), rem = mv & ((half << 1) - 1ull);
    long long k = (long long)(mv >> shift);
    if (rem > half) ++k;
    if (rem == half) irregular = true;                   // exact tie: result depends on the accumulator's parity
    return k;
}

__device__ __forceinline__ double os_from_int(long long M, int e) {   // M in [2^52, 2^53] -> M * 2^(e-52)
    if (M >= kTwo53) return __longlong_as_double((long long)(e + 1 + 1023) << 52);
    return __longlong_as_double(((long long)(e + 1023) << 52) | (M - kTwo52));
}

// ---- (B) per chunk: predicted binade and K_c
__global__ void __launch_bounds__(256) os_quantise_kernel(const double* __restrict__ x, int64_t n, int mode, const double* __restrict__ div,
                                                          double mean, const double* __restrict__ csum, const double* __restrict__ cpre,
                                                          int* __restrict__ ce, long long* __restrict__ cK, int64_t n_chunks) {
    const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const int lane = threadIdx.x & 31;
    const double p0 = cpre[c], p1 = p0 + csum[c];
    // usable prediction: finite, well inside the normal range, and the whole chunk clear of the binade's edges
    bool simple = (c > 0) && (p0 > 1e-270) && (p1 < 1e270);
    int e = 0;
    if (simple) {
        e = (int)((unsigned long long)__double_as_longlong(p0) >> 52) - 1023;
        const double lo = __longlong_as_double((long long)(e + 1023) << 52), hi = lo * 2.0;
        simple = (p0 * (1.0 - 1e-6) > lo) && (p1 * (1.0 + 1e-6) < hi);
    }
    if (!simple) {                                       // warp-uniform
        if (lane == 0) { ce[c] = kOsComplex; cK[c] = 0; }
        return;
    }
    const double dv = (mode == 1) ? *div : 1.0;
    long long K = 0;
    bool irregular = false;
#pragma unroll
    for (int k = 0; k < kOsL / 32; ++k) {
        const int64_t i = c * kOsL + k * 32 + lane;
        if (i < n) K += os_quantise(os_term(mode, x[i], dv, mean), e, irregular);
    }
    for (int o = 16; o; o >>= 1) K += __shfl_xor_sync(0xffffffffu, K, o);
    irregular = __any_sync(0xffffffffu, irregular);
    if (lane == 0) { ce[c] = irregular ? kOsComplex : e; cK[c] = K; }
}

// ---- (C) exact carry propagation, one warp.  32 chunks per step: within a run of simple chunks that
// share the accumulator's binade the recurrence is an integer prefix sum (m_j = m + K_0 + ... + K_j, all
// K >= 0), so the warp scans the 32 K's with shuffles, finds the first chunk that breaks the run (complex
// flag, other predicted binade, or m_j > 2^53) with one ballot, commits every chunk before it at once and
// replays only that chunk with the scalar chain.
__global__ void __launch_bounds__(32) os_carry_kernel(double* __restrict__ io, int64_t n, int mode, const double* __restrict__ div, double mean,
                                                      const int* __restrict__ ce, const long long* __restrict__ cK, int64_t n_chunks,
                                                      double* __restrict__ carry_in, unsigned char* __restrict__ replayed,
                                                      double* __restrict__ out) {
    __shared__ double vbuf[kOsL];
    const int lane = threadIdx.x;
    const double dv = (mode == 1) ? *div : 1.0;
    double s = 0.0;                                      // uniform across the warp
    int64_t c = 0;
    while (c < n_chunks) {
        const int64_t j = c + lane;
        const bool in = j < n_chunks;
        const int e_j = in ? ce[j] : kOsComplex;
        const long long K_j = in ? cK[j] : 0;
        const long long sb = __double_as_longlong(s);
        const int e_s = (int)((unsigned long long)sb >> 52) - 1023;      // s <= 0, NaN, inf never match a predicted binade
        const long long m = (sb & 0xFFFFFFFFFFFFFll) | kTwo52;
        long long inc = K_j;                                            // inclusive prefix of K over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        // a chunk with K_j > 2^53 would overflow nothing here (sums of 32 x 2^61 fit) but fails the bound below
        // carry-in still inside the binade (a chunk may END exactly on 2^(e+1); its successor then sees the new
        // binade on the next step) and carry-out not beyond it
        const bool ok = in && e_j != kOsComplex && e_j == e_s && (m + inc - K_j) < kTwo53 && (m + inc) <= kTwo53 && inc >= 0;
        const unsigned bad = ~__ballot_sync(0xffffffffu, ok);
        const int run = bad ? (__ffs(bad) - 1) : 32;                    // chunks c .. c+run-1 are committed
        if (lane < run) carry_in[j] = os_from_int(m + inc - K_j, e_s);
        if (run > 0) {
            const long long tot = __shfl_sync(0xffffffffu, inc, run - 1);
            s = os_from_int(m + tot, e_s);
        }
        c += run;
        if (run < 32 && c < n_chunks) {                  // replay chunk c with the scalar chain
            const int64_t j0 = c * kOsL;
            const int n_el = (int)min((int64_t)kOsL, n - j0);
#pragma unroll
            for (int k = 0; k < kOsL / 32; ++k) {
                const int idx = k * 32 + lane;
                vbuf[idx] = idx < n_el ? os_term(mode, io[j0 + idx], dv, mean) : 0.0;
            }
            __syncwarp();
            if (lane == 0) {
                carry_in[c] = s;
                replayed[c] = 1;
                int q = 0;
                if (j0 == 0) { s = vbuf[0]; q = 1; }
#pragma unroll 8
                for (; q < n_el; ++q) {
                    s = __dadd_rn(s, vbuf[q]);
                    if (mode == 1) vbuf[q] = s;
                }
            }
            __syncwarp();
            if (mode == 1) {
#pragma unroll
                for (int k = 0; k < kOsL / 32; ++k) {
                    const int idx = k * 32 + lane;
                    if (idx < n_el) io[j0 + idx] = vbuf[idx];
                }
            }
            __syncwarp();
            s = __shfl_sync(0xffffffffu, s, 0);
            ++c;
        }
    }
    if (lane == 0 && out) out[0] = s;
}

// ---- (D) prefix mode: per-element prefixes of the chunks that were not replayed
__global__ void __launch_bounds__(256) os_expand_kernel(double* __restrict__ io, int64_t n, const double* __restrict__ div,
                                                        const int* __restrict__ ce, const double* __restrict__ carry_in,
                                                        const unsigned char* __restrict__ replayed, int64_t n_chunks) {
    const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n_chunks || replayed[c]) return;
    const int lane = threadIdx.x & 31;
    const int e = ce[c];
    const double dv = *div;
    const long long m0 = (__double_as_longlong(carry_in[c]) & 0xFFFFFFFFFFFFFll) | kTwo52;
    // lane owns 8 consecutive elements; quantise, scan inside the lane, then across lanes
    long long k[kOsL / 32];
    bool dummy = false;
    long long run = 0;
    const int64_t j0 = c * kOsL + lane * (kOsL / 32);
#pragma unroll
    for (int q = 0; q < kOsL / 32; ++q) {
        const int64_t i = j0 + q;
        run += (i < n) ? os_quantise(__ddiv_rn(io[i], dv), e, dummy) : 0;
        k[q] = run;
    }
    long long inc = run;
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    const long long before = m0 + inc - run;
#pragma unroll
    for (int q = 0; q < kOsL / 32; ++q) {
        const int64_t i = j0 + q;
        if (i < n) io[i] = os_from_int(before + k[q], e);
    }
}

int device_ordered_sum(double* d_x, int64_t n, int mode, const double* d_div, double mean, double* d_out, cudaStream_t s) {
    static const bool force_serial = getenv("OTSLAM_ORDERED_SERIAL") != nullptr;       // dev switch for A/B timing
    if (n < 8 * kOsL || force_serial) {
        ordered_accumulate_kernel<<<1, 32, 0, s>>>(d_x, n, mode, d_div, mean, d_out);
        OT_LAUNCHED();
        return OTSLAM_OK;
    }
    const int64_t nc = (n + kOsL - 1) / kOsL;
    // layout: csum f64 | cpre f64 | carry f64 | cK i64 | ce i32 | replayed u8
    const size_t need = (size_t)nc * (8 + 8 + 8 + 8 + 4 + 1) + 64;
    DevBuf<unsigned char> scratch;   // back to the scratch cache at exit; later users are ordered behind us on the stream
    OT_CUDA(scratch.alloc(need));
    double* csum = reinterpret_cast<double*>(scratch.p);
    double* cpre = csum + nc;
    double* carry = cpre + nc;
    long long* cK = reinterpret_cast<long long*>(carry + nc);
    int* ce = reinterpret_cast<int*>(cK + nc);
    unsigned char* replayed = reinterpret_cast<unsigned char*>(ce + nc);
    OT_CUDA(cudaMemsetAsync(replayed, 0, (size_t)nc, s));
    const unsigned grid = (unsigned)((nc + 7) / 8);
    os_chunk_sum_kernel<<<grid, 256, 0, s>>>(d_x, n, mode, d_div, mean, csum, nc);
    OT_LAUNCHED();
    os_prefix_kernel<<<1, 1024, 0, s>>>(csum, cpre, nc);
    OT_LAUNCHED();
    os_quantise_kernel<<<grid, 256, 0, s>>>(d_x, n, mode, d_div, mean, csum, cpre, ce, cK, nc);
    OT_LAUNCHED();
    os_carry_kernel<<<1, 32, 0, s>>>(d_x, n, mode, d_div, mean, ce, cK, nc, carry, replayed, d_out);
    OT_LAUNCHED();
    if (mode == 1) {
        os_expand_kernel<<<grid, 256, 0, s>>>(d_x, n, d_div, ce, carry, replayed, nc);
        OT_LAUNCHED();
    }
    return OTSLAM_OK;
}

// ---- self-test inputs
__global__ void __launch_bounds__(256) os_fill_kernel(double* __restrict__ x, int64_t n, uint64_t seed, int pattern) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t h = ((uint64_t)i + seed) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; h *= 0x94D049BB133111EBull; h ^= h >> 31;
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    double v;
    if (pattern == 0) v = 1e-5 * (0.25 + u);                                         // triangle areas of a 5 mm mesh
    else if (pattern == 1) v = exp2((double)((int)(h & 127) - 64)) * (1.0 + u);          // 2^-64 .. 2^64
    else if (pattern == 2) v = (double)(1 + (h & 1023)) * (1.0 / 1024.0);               // coarse grid: exact ties are frequent
    else if (pattern == 3) v = (h & 7) ? ((h & 8) ? 0.0 : 1e-3 * u) : 1e3 * u;          // zeros and jumps
    else v = (i % 100003 == 7) ? ((h & 1) ? -1.0 : INFINITY) : u;                       // terms outside [0, inf)
    x[i] = v;
}

__global__ void __launch_bounds__(256) os_compare_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                                         unsigned long long* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) atomicAdd(bad, 1ull);
}

}  // namespace otslam

using namespace otslam;

extern "C" int otslam_selftest_ordered_sum(int64_t n, uint64_t seed, int pattern, uint64_t* mismatches, int device) {
    if (!mismatches || n < 1 || pattern < 0 || pattern > 4) return set_error(OTSLAM_ERR_INVALID, "bad arguments");
    OT_TRY(use_device(device));
    DevBuf<double> x, a, b, tot;
    DevBuf<unsigned long long> bad;
    OT_CUDA(x.alloc(n)); OT_CUDA(a.alloc(n)); OT_CUDA(b.alloc(n)); OT_CUDA(tot.alloc(8)); OT_CUDA(bad.alloc(1));
    OT_CUDA(cudaMemset(bad.p, 0, 8));
    const unsigned grid = (unsigned)((n + 255) / 256);
    os_fill_kernel<<<grid, 256>>>(x.p, n, seed, pattern);
    OT_LAUNCHED();
    // totals, modes 0 / 2 / 3: scalar chain -> tot[0..2], parallel -> tot[4..6]
    const int modes[3] = {0, 2, 3};
    for (int k = 0; k < 3; ++k) {
        ordered_accumulate_kernel<<<1, 32>>>(x.p, n, modes[k], nullptr, 0.37, tot.p + k);
        OT_LAUNCHED();
        OT_TRY(device_ordered_sum(x.p, n, modes[k], nullptr, 0.37, tot.p + 4 + k, 0));
    }
    // prefix mode with the mode-0 total as divisor (pattern 4 has a NaN/inf total: use 1.0 there)
    if (pattern == 4) { const double one = 1.0; OT_CUDA(cudaMemcpy(tot.p + 3, &one, 8, cudaMemcpyHostToDevice)); }
    else OT_CUDA(cudaMemcpy(tot.p + 3, tot.p, 8, cudaMemcpyDeviceToDevice));
    OT_CUDA(cudaMemcpy(a.p, x.p, (size_t)n * 8, cudaMemcpyDeviceToDevice));
    OT_CUDA(cudaMemcpy(b.p, x.p, (size_t)n * 8, cudaMemcpyDeviceToDevice));
    ordered_accumulate_kernel<<<1, 32>>>(a.p, n, 1, tot.p + 3, 0.0, nullptr);
    OT_LAUNCHED();
    OT_TRY(device_ordered_sum(b.p, n, 1, tot.p + 3, 0.0, nullptr, 0));
    os_compare_kernel<<<grid, 256>>>(a.p, b.p, n, bad.p);
    OT_LAUNCHED();
    os_compare_kernel<<<1, 256>>>(tot.p, tot.p + 4, 3, bad.p);
    OT_LAUNCHED();
    unsigned long long h = 0;
    OT_CUDA(cudaMemcpy(&h, bad.p, 8, cudaMemcpyDeviceToHost));
    *mismatches = h;
    return OTSLAM_OK;
}
