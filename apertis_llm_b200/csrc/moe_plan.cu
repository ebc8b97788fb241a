// MoE dispatch on the device: capacity plan (bit-exact restatement of the reference's slot-major /
// expert-inner loop with keep-the-largest-gate overflow, core.py:508-511, 547-590), token permutation
// with the per-expert LayerNorm fused in (core.py:593 + :436), weighted un-permutation (core.py:605),
// their backward passes and per-expert column sums for the bias gradients.  No host synchronisation:
// everything downstream reads counts / offsets from device memory.
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ void ld4g(const T* p, float* f) {
    if constexpr (sizeof(T) == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    }
}

// ---- expert parallelism over peer memory (NVLink): rows of the permuted layout that live in ANOTHER rank's buffer.
// The local layout is [W destination ranks][rows_per_peer]; row r belongs to rank d = r / rows_per_peer and sits there in
// block `rank` (the source) of a [W sources][rows_per_peer] buffer whose peer-mapped base address is base[d].
// n == 0: plain local buffer.
constexpr int AB_MAX_PEERS = 16;
struct PeerRows {
    unsigned char* base[AB_MAX_PEERS];
    int n, rows_per_peer, rank;
};
__device__ __forceinline__ unsigned char* peer_row(const PeerRows& pr, void* local, int64_t r, int Dm, int es) {
    if (pr.n == 0) return reinterpret_cast<unsigned char*>(local) + (size_t)r * Dm * es;
    const int d = (int)(r / pr.rows_per_peer);
    const int64_t j = r - (int64_t)d * pr.rows_per_peer;
    return pr.base[d] + ((size_t)pr.rank * pr.rows_per_peer + j) * Dm * es;
}

// four consecutive elements as ONE store (8 bytes bf16 / 16 bytes fp32): whole sectors also when the row is peer memory
template <typename T>
__device__ __forceinline__ void st4(T* p, float a, float b, float c, float d) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}

constexpr int PLAN_THREADS = 1024;

// exclusive prefix of a small per-thread count over the CTA (thread order) + CTA total; two __syncthreads
__device__ __forceinline__ int block_excl_scan(int cnt, int* warp_tot /*[32]*/, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int before = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) {
        const int c = warp_tot[i];
        if (i < warp) before += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return before + inc - cnt;
}

// One CTA per expert.  row_local[s,k] = position of the kept (token, slot) inside the expert's segment, -1 if dropped.
// Per slot: (A) the expert's candidates are compacted in token order into a shared-memory list (token id, weight bits)
// while the top-byte histogram is built; (B) on overflow of the capacity the rem-th largest weight is radix-selected on
// the list; (C) positions are assigned in list (= token) order.  If the candidates do not fit the list the slot falls
// back to the same algorithm over global memory.
constexpr int TPT = 8;   // consecutive tokens per thread per compaction sweep

// From a 256-bin histogram: the bin holding the `need`-th largest element counted from the top, and how many are still needed
// inside it.  Warp 0: lane l owns bins [8l, 8l+8); suffix sums over the lanes by shuffles instead of a 256-step serial scan.
__device__ __forceinline__ void radix_select_bin(int* hist, int& need, int& sel_bin, int* s_sel_bin, int* s_need) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int h[8], mine = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { h[i] = hist[lane * 8 + i]; mine += h[i]; }
        // elements in the bins of higher lanes: inclusive suffix scan by shuffles, minus the lane's own count
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += v;
        }
        const int above = incl - mine;
        const int nd0 = need;
        const bool here = above < nd0 && above + mine >= nd0;       // the wanted element lies in this lane's bins
        const unsigned who = __ballot_sync(0xffffffffu, here);
        if (who == 0u) {                                            // fewer than `need` elements in total: lowest bin
            if (lane == 0) { *s_sel_bin = 0; *s_need = nd0 - (above + mine) + h[0]; }
        } else if (lane == __ffs(who) - 1) {
            int nd = nd0 - above, bin = 7;
            for (; bin > 0; --bin) {
                if (h[bin] >= nd) break;
                nd -= h[bin];
            }
            *s_sel_bin = lane * 8 + bin;
            *s_need = nd;
        }
    }
    __syncthreads();
    sel_bin = *s_sel_bin;
    need = *s_need;
    __syncthreads();
}

// histogram increment with one shared-memory atomic per distinct bin per warp: gate weights share their top bytes, so
// per-element atomics on the same bin would serialise.  Must be reached by all 32 lanes.
__device__ __forceinline__ void hist_add_warp(int* hist, bool valid, int bin) {
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
    }
}

__global__ void __launch_bounds__(PLAN_THREADS) plan_count_kernel(const int32_t* __restrict__ idx, const float* __restrict__ w,
                                                                  const int32_t* __restrict__ active, int cap,
                                                                  int32_t* __restrict__ counts, int32_t* __restrict__ row_local,
                                                                  int S, int K, int list_cap) {
    extern __shared__ uint2 s_list[];          // [list_cap] (token, weight bits) of the current slot's candidates
    __shared__ int hist[256];
    __shared__ int warp_tot[32];
    __shared__ int s_sel_bin, s_need;
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const bool is_active = active == nullptr || active[e] != 0;
    int kept_total = 0;
    for (int k = 0; k < K; ++k) {
        // ---- (A) compaction in token order (histograms are only built when the capacity overflows, in (B))
        int n_cand = 0;
        const int span = blockDim.x * TPT;
        for (int s0 = 0; s0 < S; s0 += span) {
            const int sb = s0 + tid * TPT;
            uint32_t wb[TPT];
            bool c[TPT];
            int mine = 0;
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                const int s = sb + t;
                c[t] = s < S && idx[(size_t)s * K + k] == e;
                wb[t] = c[t] ? __float_as_uint(w[(size_t)s * K + k]) : 0u;
                mine += c[t];
            }
            int tot;
            int pos = n_cand + block_excl_scan(mine, warp_tot, tot);
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                if (c[t]) {
                    if (pos < list_cap) s_list[pos] = make_uint2((uint32_t)(sb + t), wb[t]);
                    ++pos;
                }
            }
            n_cand += tot;
        }
        __syncthreads();
        const bool in_smem = n_cand <= list_cap;
        const int rem = cap - kept_total;
        int mode = 0;                       // 0 none, 1 all, 2 select the `rem` largest
        if (is_active && n_cand > 0 && rem > 0) mode = n_cand <= rem ? 1 : 2;
        uint32_t tau = 0;
        int n_eq_take = 0;
        if (mode == 2) {
            // ---- (B) radix select of the rem-th largest weight (positive floats order like their bit patterns)
            uint32_t prefix = 0, mask = 0;
            int need = rem, bin;
            for (int pass = 3; pass >= 0; --pass) {
                for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
                __syncthreads();
                if (in_smem) {
                    for (int i0 = 0; i0 < n_cand; i0 += blockDim.x) {
                        const int i = i0 + tid;
                        const uint32_t b = i < n_cand ? s_list[i].y : 0u;
                        hist_add_warp(hist, i < n_cand && (b & mask) == prefix, (int)((b >> (8 * pass)) & 0xff));
                    }
                } else {
                    for (int s0 = 0; s0 < S; s0 += blockDim.x) {
                        const int s = s0 + tid;
                        const bool c = s < S && idx[(size_t)s * K + k] == e;
                        const uint32_t b = c ? __float_as_uint(w[(size_t)s * K + k]) : 0u;
                        hist_add_warp(hist, c && (b & mask) == prefix, (int)((b >> (8 * pass)) & 0xff));
                    }
                }
                __syncthreads();
                radix_select_bin(hist, need, bin, &s_sel_bin, &s_need);
                prefix |= (uint32_t)bin << (8 * pass);
                mask |= 0xffu << (8 * pass);
            }
            tau = prefix;
            n_eq_take = need;
        }
        // ---- (C) positions in token order
        int run_eq = 0, run_kept = 0;
        const int n_items = in_smem ? n_cand : S;
        for (int s0 = 0; s0 < n_items; s0 += span) {
            const int ib = s0 + tid * TPT;
            bool c[TPT], gt[TPT], eq[TPT];
            int tokv[TPT];
            int n_eq = 0;
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                const int i = ib + t;
                uint32_t b = 0u;
                if (in_smem) {
                    c[t] = i < n_cand;
                    if (c[t]) { const uint2 it = s_list[i]; tokv[t] = (int)it.x; b = it.y; } else tokv[t] = 0;
                } else {
                    c[t] = i < S && idx[(size_t)i * K + k] == e;
                    tokv[t] = i;
                    if (c[t]) b = __float_as_uint(w[(size_t)i * K + k]);
                }
                gt[t] = c[t] && (mode == 1 || (mode == 2 && b > tau));
                eq[t] = c[t] && mode == 2 && b == tau;
                n_eq += eq[t];
            }
            int tot_eq = 0, tot_kept = 0;
            int eq_rank = run_eq;
            if (mode == 2) eq_rank += block_excl_scan(n_eq, warp_tot, tot_eq);
            bool kept[TPT];
            int n_kept = 0;
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                kept[t] = gt[t] || (eq[t] && eq_rank < n_eq_take);
                eq_rank += eq[t];
                n_kept += kept[t];
            }
            int pos = kept_total + run_kept + block_excl_scan(n_kept, warp_tot, tot_kept);
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                if (c[t]) row_local[(size_t)tokv[t] * K + k] = kept[t] ? pos : -1;
                pos += kept[t];
            }
            run_eq += tot_eq;
            run_kept += tot_kept;
        }
        kept_total += run_kept;
        __syncthreads();
    }
    if (tid == 0) counts[e] = kept_total;
}

__global__ void plan_finalize_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ counts,
                                     const int32_t* __restrict__ row_local, int32_t* __restrict__ seg_off,
                                     int32_t* __restrict__ row_of, int32_t* __restrict__ tok_of_row,
                                     int32_t* __restrict__ slot_of_row, int32_t* __restrict__ tile_expert,
                                     int32_t* __restrict__ n_rows, int S, int K, int E, int align, int64_t max_rows,
                                     int fixed_seg) {
    __shared__ int soff[34];
    if (threadIdx.x == 0) {
        int o = 0, kept = 0;
        for (int e = 0; e < E; ++e) {
            soff[e] = o;
            if (fixed_seg > 0 && counts[e] > fixed_seg) __trap();      // the caller sized the segments too small: never overflow into a neighbour
            o += fixed_seg > 0 ? fixed_seg : (counts[e] + align - 1) / align * align;
            kept += counts[e];
        }
        soff[E] = o;
        soff[E + 1] = kept;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i <= E; i += blockDim.x) seg_off[i] = soff[i];
        if (threadIdx.x == 0) { n_rows[0] = soff[E]; n_rows[1] = soff[E + 1]; }
        const int ntiles = (int)(max_rows / align);
        for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
            const int r = t * align;
            int ex = -1;
            for (int e = 0; e < E; ++e)
                if (r >= soff[e] && r < soff[e + 1]) ex = e;
            tile_expert[t] = ex;
        }
    }
    const int64_t n = (int64_t)S * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int rl = row_local[i];
        if (rl >= 0) {
            const int r = soff[idx[i]] + rl;
            row_of[i] = r;
            tok_of_row[r] = (int)(i / K);
            slot_of_row[r] = (int)(i % K);
        } else {
            row_of[i] = -1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// permute + LayerNorm (warp per row)
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) permute_ln_kernel(const TI* __restrict__ x, const float* __restrict__ stats,
                                                         const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                         const int32_t* __restrict__ tok_of_row,
                                                         const int32_t* __restrict__ tile_expert,
                                                         const int32_t* __restrict__ n_rows, TO* __restrict__ xn, int Dm,
                                                         int align, const PeerRows pr) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int total = n_rows[0];
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < total; r += gridDim.x * wpb) {
        const int tok = tok_of_row[r];
        TO* orow = reinterpret_cast<TO*>(peer_row(pr, xn, r, Dm, (int)sizeof(TO)));     // the owner's receive buffer under EP
        if (tok < 0) {
            for (int d = lane * 4; d < Dm; d += 128) st4<TO>(orow + d, 0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int e = tile_expert[r / align];
        const float mean = stats[2 * (size_t)tok], rstd = stats[2 * (size_t)tok + 1];
        const TI* irow = x + (size_t)tok * Dm;
        const float* g = ln_w + (size_t)e * Dm;
        const float* bb = ln_b + (size_t)e * Dm;
        for (int d = lane * 4; d < Dm; d += 128) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g + d));
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bb + d));
            float xv[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) xv[v] = ab_to_float(irow[d + v]);
            const float o0 = fmaf((xv[0] - mean) * rstd, gv.x, bv.x), o1 = fmaf((xv[1] - mean) * rstd, gv.y, bv.y);
            const float o2 = fmaf((xv[2] - mean) * rstd, gv.z, bv.z), o3 = fmaf((xv[3] - mean) * rstd, gv.w, bv.w);
            st4<TO>(orow + d, o0, o1, o2, o3);
        }
    }
}

// out[s,:] = sum_k w[s,k] * y[row_of[s,k],:]   (warp per token, fixed slot order).  The loads of a group of chunks - 4
// elements of every one of the K rows each - are all issued before the first is used.
template <typename TY, typename TO, int KM>
__global__ void __launch_bounds__(256) unpermute_kernel(const TY* __restrict__ y, const int32_t* __restrict__ row_of,
                                                        const float* __restrict__ w, const float* __restrict__ res, TO* __restrict__ out,
                                                        const uint32_t* __restrict__ seed, uint32_t thresh, float scale, int S, int K, int Dm) {
    constexpr int DU = 8 / KM;           // chunks in flight (KM = upper bound of the experts per token)
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    uint32_t s0 = 0, s1 = 0;
    if (seed) { s0 = __ldg(seed); s1 = __ldg(seed + 1); }
    for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
        TO* orow = out + (size_t)s * Dm;
        const TY* rows[KM];
        float wk[KM];
        int rid[KM];
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            rid[k] = k < K ? row_of[(size_t)s * K + k] : -1;
            wk[k] = rid[k] >= 0 ? w[(size_t)s * K + k] : 0.f;
            rows[k] = rid[k] >= 0 ? y + (size_t)rid[k] * Dm : nullptr;
        }
        for (int d0 = lane * 4; d0 < Dm; d0 += 128 * DU) {
            float yv[DU][KM][4];
#pragma unroll
            for (int u = 0; u < DU; ++u) {
                const int d = d0 + u * 128;
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                    if (k < K && rows[k] != nullptr && d < Dm) ld4g<TY>(rows[k] + d, yv[u][k]);
                    else { yv[u][k][0] = yv[u][k][1] = yv[u][k][2] = yv[u][k][3] = 0.f; }
                }
            }
#pragma unroll
            for (int u = 0; u < DU; ++u) {
                const int d = d0 + u * 128;
                if (d >= Dm) break;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                    if (k < K && rows[k] != nullptr) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) acc[v] += yv[u][k][v] * wk[k];   // separate mul and add, as index_add_(y*w)
                    }
                }
                // the caller's output dropout and residual add (core.py:918-919) in the same pass
                if (seed) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[v] = ab_out_keep(s0, s1, (uint64_t)s * Dm + d + v, thresh) ? acc[v] * scale : 0.f;
                }
                if (res) {
                    const float4 rr = *reinterpret_cast<const float4*>(res + (size_t)s * Dm + d);
                    acc[0] += rr.x; acc[1] += rr.y; acc[2] += rr.z; acc[3] += rr.w;
                }
                st4<TO>(orow + d, acc[0], acc[1], acc[2], acc[3]);
            }
        }
    }
}

// dy[r,:] = w[r] * dout[tok,:] ; dw_row[r] = <dout[tok,:], y[r,:]>
template <typename TD, typename TY, typename TO>
__global__ void __launch_bounds__(256) unpermute_bwd_kernel(const TD* __restrict__ dout, const TY* __restrict__ y,
                                                            const float* __restrict__ w, const int32_t* __restrict__ tok_of_row,
                                                            const int32_t* __restrict__ slot_of_row,
                                                            const int32_t* __restrict__ n_rows, TO* __restrict__ dy,
                                                            float* __restrict__ dw_row, const uint32_t* __restrict__ seed, uint32_t thresh,
                                                            float scale, int K, int Dm, const PeerRows pr) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int total = n_rows[0];
    uint32_t s0 = 0, s1 = 0;
    if (seed) { s0 = __ldg(seed); s1 = __ldg(seed + 1); }
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < total; r += gridDim.x * wpb) {
        const int tok = tok_of_row[r];
        TO* orow = reinterpret_cast<TO*>(peer_row(pr, dy, r, Dm, (int)sizeof(TO)));     // the owner's receive buffer under EP
        if (tok < 0) {
            for (int d = lane * 4; d < Dm; d += 128) st4<TO>(orow + d, 0.f, 0.f, 0.f, 0.f);
            if (lane == 0) dw_row[r] = 0.f;
            continue;
        }
        const float wk = w[(size_t)tok * K + slot_of_row[r]];
        const TD* drow = dout + (size_t)tok * Dm;
        const TY* yrow = y + (size_t)r * Dm;
        float dot = 0.f;
        for (int d = lane * 4; d < Dm; d += 128) {
            float o[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float g = ab_to_float(drow[d + v]);
                if (seed) g = ab_out_keep(s0, s1, (uint64_t)tok * Dm + d + v, thresh) ? g * scale : 0.f;      // the forward's output-dropout mask
                dot = fmaf(g, ab_to_float(yrow[d + v]), dot);
                o[v] = g * wk;
            }
            st4<TO>(orow + d, o[0], o[1], o[2], o[3]);
        }
        dot = ab_warp_sum(dot);
        if (lane == 0) dw_row[r] = dot;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward per permuted row (warp per row) and per-tile partial sums of the affine grads
// ---------------------------------------------------------------------------------------------
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) permute_ln_bwd_rows_kernel(const TG* __restrict__ dxn, const TX* __restrict__ x,
                                                                  const float* __restrict__ stats, const float* __restrict__ ln_w,
                                                                  const int32_t* __restrict__ tok_of_row,
                                                                  const int32_t* __restrict__ tile_expert,
                                                                  const int32_t* __restrict__ n_rows, float* __restrict__ dxrow,
                                                                  int Dm, int align) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int total = n_rows[0];
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < total; r += gridDim.x * wpb) {
        const int tok = tok_of_row[r];
        if (tok < 0) continue;
        const int e = tile_expert[r / align];
        const float* g = ln_w + (size_t)e * Dm;
        const float mean = stats[2 * (size_t)tok], rstd = stats[2 * (size_t)tok + 1];
        const TX* xr = x + (size_t)tok * Dm;
        const TG* gr = dxn + (size_t)r * Dm;
        float a1 = 0.f, a2 = 0.f;
        for (int d = lane * 4; d < Dm; d += 128) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g + d));
            const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float dh = ab_to_float(gr[d + v]) * gg[v];
                const float xh = (ab_to_float(xr[d + v]) - mean) * rstd;
                a1 += dh;
                a2 = fmaf(dh, xh, a2);
            }
        }
        const float m1 = ab_warp_sum(a1) / (float)Dm, m2 = ab_warp_sum(a2) / (float)Dm;
        float* orow = dxrow + (size_t)r * Dm;
        for (int d = lane * 4; d < Dm; d += 128) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g + d));
            const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
            float o[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float dh = ab_to_float(gr[d + v]) * gg[v];
                const float xh = (ab_to_float(xr[d + v]) - mean) * rstd;
                o[v] = rstd * (dh - m1 - xh * m2);
            }
            *reinterpret_cast<float4*>(orow + d) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Row pass and per-tile column partials in one kernel (hidden sizes up to NIT * 128): one CTA per 128-row tile (one
// expert), a lane always owns the same columns, so dxn * xhat and dxn accumulate in registers while the rows stream; the
// warps then add their sums in warp order.  Replaces the rows + cols kernel pair (one read of dxn and x instead of two).
template <typename TX, typename TG, int NIT>
__global__ void __launch_bounds__(256, NIT <= 6 ? 3 : 2) permute_ln_bwd_fused_kernel(const TG* __restrict__ dxn, const TX* __restrict__ x,
                                                                   const float* __restrict__ stats, const float* __restrict__ ln_w,
                                                                   const int32_t* __restrict__ tok_of_row,
                                                                   const int32_t* __restrict__ tile_expert,
                                                                   const int32_t* __restrict__ n_rows, float* __restrict__ dxrow,
                                                                   float* __restrict__ part, int Dm, int align, int ratio) {
    // `align` = rows of this kernel's work tile, `ratio` work tiles per tile_expert entry (the GEMM's row tile)
    __shared__ float sacc[2][NIT * 128];
    const int t = blockIdx.x;
    if ((int64_t)t * align >= n_rows[0]) return;
    const int e = tile_expert[t / ratio];
    if (e < 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const float* g = ln_w + (size_t)e * Dm;
    float gw[NIT][4], gb[NIT][4];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { gw[it][v] = 0.f; gb[it][v] = 0.f; }
    }
    // The loads of a row are issued piecewise (register budget), each piece a full memory latency when it misses: the
    // NEXT row of this warp (token id looked up two rows ahead) is therefore prefetched into L1 while this one is processed.
    int tok_cur = tok_of_row[t * align + warp];
    int tok_nxt = warp + wpb < align ? tok_of_row[t * align + warp + wpb] : -1;
    for (int i = warp; i < align; i += wpb) {
        const int r = t * align + i;
        const int tok = tok_cur;
        tok_cur = tok_nxt;
        tok_nxt = i + 2 * wpb < align ? tok_of_row[t * align + i + 2 * wpb] : -1;
        if (tok_cur >= 0) {
            const char* px = reinterpret_cast<const char*>(x + (size_t)tok_cur * Dm);
            const char* pg = reinterpret_cast<const char*>(dxn + (size_t)(r + wpb) * Dm);
            for (int o = lane * 128; o < Dm * (int)sizeof(TX); o += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(px + o));
            for (int o = lane * 128; o < Dm * (int)sizeof(TG); o += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pg + o));
        }
        if (tok < 0) continue;
        const float mean = stats[2 * (size_t)tok], rstd = stats[2 * (size_t)tok + 1];
        const TX* xr = x + (size_t)tok * Dm;
        const TG* gr = dxn + (size_t)r * Dm;
        // pass 1: column partials and the two row sums (the row is read again in pass 2 from L1: keeping it in
        // registers would halve the resident warps)
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int d = it * 128 + lane * 4;
            if (d < Dm) {
                float fx[4], fg[4], gg[4];
                ld4g<TX>(xr + d, fx);
                ld4g<TG>(gr + d, fg);
                ld4g<float>(g + d, gg);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float xh = (fx[v] - mean) * rstd;
                    gw[it][v] = fmaf(fg[v], xh, gw[it][v]);
                    gb[it][v] += fg[v];
                    const float dh = fg[v] * gg[v];
                    a1 += dh;
                    a2 = fmaf(dh, xh, a2);
                }
            }
        }
        const float m1 = ab_warp_sum(a1) / (float)Dm, m2 = ab_warp_sum(a2) / (float)Dm;
        float* orow = dxrow + (size_t)r * Dm;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int d = it * 128 + lane * 4;
            if (d < Dm) {
                float fx[4], fg[4], gg[4], o[4];
                ld4g<TX>(xr + d, fx);
                ld4g<TG>(gr + d, fg);
                ld4g<float>(g + d, gg);
#pragma unroll
                for (int v = 0; v < 4; ++v) o[v] = rstd * (fg[v] * gg[v] - m1 - (fx[v] - mean) * rstd * m2);
                *reinterpret_cast<float4*>(orow + d) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    for (int wv = 0; wv < wpb; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int i = it * 128 + lane * 4 + v;
                    sacc[0][i] = wv ? sacc[0][i] + gw[it][v] : gw[it][v];
                    sacc[1][i] = wv ? sacc[1][i] + gb[it][v] : gb[it][v];
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 2 * Dm; i += blockDim.x) {
        const int q = i / Dm, d = i % Dm;
        part[((size_t)t * 2 + q) * Dm + d] = sacc[q][d];
    }
}

// grid (tiles, ceil(Dm/128)); block (32 lanes x 4 columns, 8 warps over the tile's rows)
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) permute_ln_bwd_cols_kernel(const TG* __restrict__ dxn, const TX* __restrict__ x,
                                                                  const float* __restrict__ stats,
                                                                  const int32_t* __restrict__ tok_of_row,
                                                                  const int32_t* __restrict__ tile_expert,
                                                                  const int32_t* __restrict__ n_rows, float* __restrict__ part,
                                                                  int Dm, int align, int ratio) {
    __shared__ float red[8][2][128];
    const int t = blockIdx.x;
    if ((int64_t)t * align >= n_rows[0] || tile_expert[t / ratio] < 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.y * 128 + lane * 4;
    float gw[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
    if (d < Dm) {
        for (int i = warp; i < align; i += 8) {
            const int r = t * align + i;
            const int tok = tok_of_row[r];
            if (tok < 0) continue;
            const float mean = stats[2 * (size_t)tok], rstd = stats[2 * (size_t)tok + 1];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float dv = ab_to_float(dxn[(size_t)r * Dm + d + v]);
                const float xh = (ab_to_float(x[(size_t)tok * Dm + d + v]) - mean) * rstd;
                gw[v] = fmaf(dv, xh, gw[v]);
                gb[v] += dv;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) { red[warp][0][lane * 4 + v] = gw[v]; red[warp][1][lane * 4 + v] = gb[v]; }
    __syncthreads();
    const int q = threadIdx.x >> 7, c = threadIdx.x & 127;      // 256 threads: (w|b) x 128 columns
    float s2 = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s2 += red[wv][q][c];
    const int dd = blockIdx.y * 128 + c;
    if (dd < Dm) part[((size_t)t * 2 + q) * Dm + dd] = s2;
}

// per-tile column sums: grid (tiles, ceil(C/256)); block (32 lanes x 8 columns, 8 warps over the rows)
template <typename T>
__global__ void __launch_bounds__(256) tile_colsum_kernel(const T* __restrict__ a, const int32_t* __restrict__ tile_expert,
                                                          const int32_t* __restrict__ n_rows, float* __restrict__ part, int C,
                                                          int align, int ratio) {
    __shared__ float red[8][256];
    const int t = blockIdx.x;
    if ((int64_t)t * align >= n_rows[0] || tile_expert[t / ratio] < 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.y * 256 + lane * 8;
    float acc[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] = 0.f;
    if (c0 < C) {
        // 8 columns per thread as whole 16-byte vectors, 4 rows in flight
        constexpr int VN = 16 / (int)sizeof(T);          // elements per vector (8 bf16 / 4 fp32)
        for (int i0 = warp; i0 < align; i0 += 32) {
            uint4 q[4][8 / VN];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 8;
                const uint4* row = reinterpret_cast<const uint4*>(a + ((size_t)t * align + i) * C + c0);
#pragma unroll
                for (int w = 0; w < 8 / VN; ++w) q[u][w] = i < align ? __ldg(row + w) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int w = 0; w < 8 / VN; ++w) {
                    float f[VN];
                    ab_vec16<T>::unpack(q[u][w], f);
#pragma unroll
                    for (int v = 0; v < VN; ++v) acc[w * VN + v] += f[v];
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) red[warp][lane * 8 + v] = acc[v];
    __syncthreads();
    const int c = threadIdx.x;
    float s2 = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s2 += red[wv][c];
    if (blockIdx.y * 256 + c < C) part[(size_t)t * C + blockIdx.y * 256 + c] = s2;
}

// out[e][j] = sum over the (contiguous) tiles of expert e of part[t][j], fixed order.  Columns [0, split) go to
// out_a [E][split], columns [split, ncols) to out_b [E][ncols-split].
__global__ void tile_reduce_kernel(const float* __restrict__ part, const int32_t* __restrict__ tile_expert,
                                   const int32_t* __restrict__ n_rows, float* __restrict__ out_a, float* __restrict__ out_b,
                                   int split, int ncols, int align, int ratio) {
    __shared__ int s_t0, s_t1;
    __shared__ int s_list[1024];                       // tile ids of this expert, in order (the common case fits)
    __shared__ int s_n;
    const int e = blockIdx.y;
    const int ntiles = n_rows[0] / align;
    if (threadIdx.x < 32) {                            // one warp builds the ordered list with ballots
        int n = 0, t0 = ntiles, t1 = 0;
        for (int base = 0; base < ntiles; base += 32) {
            const int t = base + (int)threadIdx.x;
            const bool mine = t < ntiles && tile_expert[t / ratio] == e;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                const int pos = n + __popc(m & ((1u << threadIdx.x) - 1u));
                if (pos < 1024) s_list[pos] = t;
            }
            if (m) { t0 = min(t0, base + __ffs(m) - 1); t1 = base + 32 - __clz(m); }
            n += __popc(m);
        }
        if (threadIdx.x == 0) { s_t0 = t0; s_t1 = t1; s_n = n; }
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    float s = 0.f;
    if (s_n <= 1024) {
        // loads of 8 tiles are issued before the 8 dependent adds; order = tile order -> deterministic
        for (int i0 = 0; i0 < s_n; i0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = i0 + u < s_n ? part[(size_t)s_list[i0 + u] * ncols + j] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
    } else {
        for (int t = s_t0; t < s_t1; ++t)
            if (tile_expert[t / ratio] == e) s += part[(size_t)t * ncols + j];  // tiles of an expert may be interleaved (EP layout)
    }
    if (j < split) out_a[(size_t)e * split + j] = s;
    else out_b[(size_t)e * (ncols - split) + (j - split)] = s;
}

// The per-tile reduction kernels work on tiles of at most 128 rows whatever the GEMM's row tile is (more, smaller CTAs:
// these kernels are latency-bound streams); tile_expert is looked up through the ratio of the two.
int work_rows(int row_align) { return row_align % 128 == 0 ? 128 : row_align; }

int make_peers(PeerRows* pr, const uint64_t* peer_ptrs, int W, int rank, int64_t rows_per_peer) {
    memset(pr, 0, sizeof(*pr));
    if (peer_ptrs == nullptr) return AB_OK;
    AB_REQUIRE(W >= 1 && W <= AB_MAX_PEERS && rank >= 0 && rank < W && rows_per_peer > 0 && rows_per_peer < (1ll << 31),
               "ep: bad peer table (W=%d, rank=%d, rows_per_peer=%lld; at most %d ranks)", W, rank, (long long)rows_per_peer, AB_MAX_PEERS);
    for (int i = 0; i < W; ++i) {
        AB_REQUIRE(peer_ptrs[i] != 0 && peer_ptrs[i] % 16 == 0, "ep: peer buffer %d is null or not 16-byte aligned", i);
        pr->base[i] = reinterpret_cast<unsigned char*>(peer_ptrs[i]);
    }
    pr->n = W; pr->rows_per_peer = (int)rows_per_peer; pr->rank = rank;
    return AB_OK;
}

int rows_grid(int64_t max_rows) {
    const int64_t want = ab_ceil_div(max_rows, 8);
    const int64_t cap = (int64_t)ab_num_sms() * 8;
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int64_t ab_moe_max_rows(int S, int K, int E, int cap, int row_align) {
    int64_t kept = (int64_t)S * K;
    if ((int64_t)E * cap < kept) kept = (int64_t)E * cap;
    return ab_round_up(kept + (int64_t)E * (row_align - 1), row_align);
}

extern "C" size_t ab_moe_plan_workspace_bytes(int S, int K, int E) {
    (void)E;
    return (size_t)ab_round_up((int64_t)S * K * sizeof(int32_t), 256);
}

extern "C" int ab_moe_plan(const int32_t* idx, const float* w, const int32_t* active, int cap, int32_t* counts,
                           int32_t* seg_off, int32_t* row_of, int32_t* tok_of_row, int32_t* slot_of_row,
                           int32_t* tile_expert, int32_t* n_rows, void* ws, size_t ws_bytes, int S, int K, int E,
                           int row_align, int64_t max_rows, int fixed_seg, cudaStream_t stream) {
    AB_REQUIRE(S > 0 && K >= 1 && E >= 1 && E <= 32, "moe_plan: bad shape S=%d K=%d E=%d", S, K, E);
    if (fixed_seg > 0) {
        // with a capacity limit a segment must hold `cap` rows; without one (cap >= S) the caller sizes it from the counts
        // it has reduced over the ranks, and the kernel traps rather than overflow a segment
        AB_REQUIRE(fixed_seg % row_align == 0 && (cap >= S || fixed_seg >= cap) && max_rows == (int64_t)E * fixed_seg,
                   "moe_plan: fixed_seg (%d) must be a multiple of %d, >= cap, and max_rows == E*fixed_seg", fixed_seg, row_align);
    } else {
        AB_REQUIRE(row_align > 0 && max_rows % row_align == 0 && max_rows >= ab_moe_max_rows(S, K, E, cap < S ? cap : S, row_align),
                   "moe_plan: max_rows %lld too small or not a multiple of %d", (long long)max_rows, row_align);
    }
    AB_REQUIRE(ws && ws_bytes >= ab_moe_plan_workspace_bytes(S, K, E), "moe_plan: workspace too small");
    int32_t* row_local = (int32_t*)ws;
    AB_CHECK_CUDA(cudaMemsetAsync(tok_of_row, 0xFF, (size_t)max_rows * sizeof(int32_t), stream));
    AB_CHECK_CUDA(cudaMemsetAsync(slot_of_row, 0xFF, (size_t)max_rows * sizeof(int32_t), stream));
    // candidate list in shared memory: as many entries as fit (8 B each), at most S
    int list_cap = (200 * 1024) / 8;
    if (list_cap > S) list_cap = S;
    const size_t plan_smem = (size_t)list_cap * sizeof(uint2);
    AB_CHECK_CUDA(cudaFuncSetAttribute(plan_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem));
    plan_count_kernel<<<E, PLAN_THREADS, plan_smem, stream>>>(idx, w, active, cap, counts, row_local, S, K, list_cap);
    AB_LAUNCH_CHECK();
    const int64_t want = ab_ceil_div((int64_t)S * K, 256);
    const int grid = (int)(want < ab_num_sms() * 4 ? want : ab_num_sms() * 4);
    plan_finalize_kernel<<<grid, 256, 0, stream>>>(idx, counts, row_local, seg_off, row_of, tok_of_row, slot_of_row, tile_expert,
                                                   n_rows, S, K, E, row_align, max_rows, fixed_seg);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

namespace {
int permute_ln_impl(const void* x, const float* stats, const float* ln_w, const float* ln_b, const int32_t* tok_of_row,
                    const int32_t* tile_expert, const int32_t* n_rows, void* xn, int Dm, int row_align, int64_t max_rows, int dtype,
                    int out_dtype, const PeerRows& pr, cudaStream_t stream) {
    AB_REQUIRE(Dm > 0 && Dm % 4 == 0, "moe_permute_ln: hidden size must be a multiple of 4");
    const int grid = rows_grid(max_rows);
#define AB_PLN(TI, TO) permute_ln_kernel<TI, TO><<<grid, 256, 0, stream>>>((const TI*)x, stats, ln_w, ln_b, tok_of_row, tile_expert, n_rows, (TO*)xn, Dm, row_align, pr)
    if (dtype == AB_F32 && out_dtype == AB_F32) AB_PLN(float, float);
    else if (dtype == AB_F32 && out_dtype == AB_BF16) AB_PLN(float, __nv_bfloat16);
    else if (dtype == AB_BF16 && out_dtype == AB_BF16) AB_PLN(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == AB_BF16 && out_dtype == AB_F32) AB_PLN(__nv_bfloat16, float);
    else AB_REQUIRE(false, "moe_permute_ln: bad dtypes");
#undef AB_PLN
    AB_LAUNCH_CHECK();
    return AB_OK;
}
}  // namespace

extern "C" int ab_moe_permute_ln(const void* x, const float* stats, const float* ln_w, const float* ln_b,
                                 const int32_t* tok_of_row, const int32_t* tile_expert, const int32_t* n_rows, void* xn, int Dm,
                                 int row_align, int64_t max_rows, int dtype, int out_dtype, cudaStream_t stream) {
    PeerRows pr;
    make_peers(&pr, nullptr, 0, 0, 0);
    return permute_ln_impl(x, stats, ln_w, ln_b, tok_of_row, tile_expert, n_rows, xn, Dm, row_align, max_rows, dtype, out_dtype, pr, stream);
}

extern "C" int ab_ep_permute_ln(const void* x, const float* stats, const float* ln_w, const float* ln_b,
                                const int32_t* tok_of_row, const int32_t* tile_expert, const int32_t* n_rows,
                                const uint64_t* peer_xn, int W, int rank, int64_t rows_per_peer, int Dm, int row_align,
                                int64_t max_rows, int dtype, int out_dtype, cudaStream_t stream) {
    AB_REQUIRE(peer_xn != nullptr && max_rows == (int64_t)W * rows_per_peer, "ep_permute_ln: max_rows must equal W * rows_per_peer");
    PeerRows pr;
    if (int e = make_peers(&pr, peer_xn, W, rank, rows_per_peer)) return e;
    return permute_ln_impl(x, stats, ln_w, ln_b, tok_of_row, tile_expert, n_rows, nullptr, Dm, row_align, max_rows, dtype, out_dtype, pr, stream);
}

extern "C" int ab_moe_unpermute(const void* y, const int32_t* row_of, const float* w, const float* res, void* out, float drop_p,
                                const uint32_t* drop_seed, int S, int K, int Dm, int y_dtype, int out_dtype, cudaStream_t stream) {
    AB_REQUIRE(Dm > 0 && Dm % 4 == 0, "moe_unpermute: hidden size must be a multiple of 4");
    AB_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed), "moe_unpermute: dropout p must be in [0,1) and needs a seed when > 0");
    AB_REQUIRE(res == nullptr || ((uintptr_t)res % 16) == 0, "moe_unpermute: the residual must be 16-byte aligned fp32");
    const uint32_t thresh = (uint32_t)((double)drop_p * 4294967296.0);
    const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t* sd = drop_p > 0.f ? drop_seed : nullptr;
    const int grid = rows_grid(S);
#define AB_UNP(TY, TO)                                                                                                          \
    {                                                                                                                            \
        if (K <= 2) unpermute_kernel<TY, TO, 2><<<grid, 256, 0, stream>>>((const TY*)y, row_of, w, res, (TO*)out, sd, thresh, scale, S, K, Dm); \
        else if (K <= 4) unpermute_kernel<TY, TO, 4><<<grid, 256, 0, stream>>>((const TY*)y, row_of, w, res, (TO*)out, sd, thresh, scale, S, K, Dm); \
        else unpermute_kernel<TY, TO, 8><<<grid, 256, 0, stream>>>((const TY*)y, row_of, w, res, (TO*)out, sd, thresh, scale, S, K, Dm); \
    }
    AB_REQUIRE(K >= 1 && K <= 8, "moe_unpermute: experts_per_token must be in [1, 8]");
    if (y_dtype == AB_F32 && out_dtype == AB_F32) AB_UNP(float, float)
    else if (y_dtype == AB_BF16 && out_dtype == AB_F32) AB_UNP(__nv_bfloat16, float)
    else if (y_dtype == AB_BF16 && out_dtype == AB_BF16) AB_UNP(__nv_bfloat16, __nv_bfloat16)
    else if (y_dtype == AB_F32 && out_dtype == AB_BF16) AB_UNP(float, __nv_bfloat16)
    else AB_REQUIRE(false, "moe_unpermute: bad dtypes");
#undef AB_UNP
    AB_LAUNCH_CHECK();
    return AB_OK;
}

namespace {
int unpermute_bwd_impl(const void* dout, const void* y, const float* w, const int32_t* tok_of_row, const int32_t* slot_of_row,
                       const int32_t* n_rows, void* dy, float* dw_row, float drop_p, const uint32_t* drop_seed, int K, int Dm,
                       int64_t max_rows, int dout_dtype, int y_dtype, int dy_dtype, const PeerRows& pr, cudaStream_t stream) {
    AB_REQUIRE(Dm > 0 && Dm % 4 == 0, "moe_unpermute_bwd: hidden size must be a multiple of 4");
    AB_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed), "moe_unpermute_bwd: dropout p must be in [0,1) and needs a seed when > 0");
    const uint32_t thresh = (uint32_t)((double)drop_p * 4294967296.0);
    const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t* sd = drop_p > 0.f ? drop_seed : nullptr;
    const int grid = rows_grid(max_rows);
#define AB_UB(TD, TY, TO) unpermute_bwd_kernel<TD, TY, TO><<<grid, 256, 0, stream>>>((const TD*)dout, (const TY*)y, w, tok_of_row, slot_of_row, n_rows, (TO*)dy, dw_row, sd, thresh, scale, K, Dm, pr)
    const int key = dout_dtype * 4 + y_dtype * 2 + dy_dtype;
    switch (key) {
        case 0: AB_UB(float, float, float); break;
        case 1: AB_UB(float, float, __nv_bfloat16); break;
        case 2: AB_UB(float, __nv_bfloat16, float); break;
        case 3: AB_UB(float, __nv_bfloat16, __nv_bfloat16); break;
        case 4: AB_UB(__nv_bfloat16, float, float); break;
        case 5: AB_UB(__nv_bfloat16, float, __nv_bfloat16); break;
        case 6: AB_UB(__nv_bfloat16, __nv_bfloat16, float); break;
        case 7: AB_UB(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16); break;
        default: AB_REQUIRE(false, "moe_unpermute_bwd: bad dtypes");
    }
#undef AB_UB
    AB_LAUNCH_CHECK();
    return AB_OK;
}
}  // namespace

extern "C" int ab_moe_unpermute_bwd(const void* dout, const void* y, const float* w, const int32_t* tok_of_row,
                                    const int32_t* slot_of_row, const int32_t* n_rows, void* dy, float* dw_row, float drop_p,
                                    const uint32_t* drop_seed, int K, int Dm, int64_t max_rows, int dout_dtype, int y_dtype,
                                    int dy_dtype, cudaStream_t stream) {
    PeerRows pr;
    make_peers(&pr, nullptr, 0, 0, 0);
    return unpermute_bwd_impl(dout, y, w, tok_of_row, slot_of_row, n_rows, dy, dw_row, drop_p, drop_seed, K, Dm, max_rows, dout_dtype,
                              y_dtype, dy_dtype, pr, stream);
}

extern "C" int ab_ep_unpermute_bwd(const void* dout, const void* y, const float* w, const int32_t* tok_of_row,
                                   const int32_t* slot_of_row, const int32_t* n_rows, const uint64_t* peer_dy, int W, int rank,
                                   int64_t rows_per_peer, float* dw_row, float drop_p, const uint32_t* drop_seed, int K, int Dm,
                                   int64_t max_rows, int dout_dtype, int y_dtype, int dy_dtype, cudaStream_t stream) {
    AB_REQUIRE(peer_dy != nullptr && max_rows == (int64_t)W * rows_per_peer, "ep_unpermute_bwd: max_rows must equal W * rows_per_peer");
    PeerRows pr;
    if (int e = make_peers(&pr, peer_dy, W, rank, rows_per_peer)) return e;
    return unpermute_bwd_impl(dout, y, w, tok_of_row, slot_of_row, n_rows, nullptr, dw_row, drop_p, drop_seed, K, Dm, max_rows,
                              dout_dtype, y_dtype, dy_dtype, pr, stream);
}

extern "C" size_t ab_moe_permute_ln_bwd_workspace_bytes(int Dm, int row_align, int64_t max_rows) {
    return (size_t)ab_round_up((max_rows / work_rows(row_align)) * 2 * (int64_t)Dm * sizeof(float), 256);
}

extern "C" int ab_moe_permute_ln_bwd(const void* dxn, const void* x, const float* stats, const float* ln_w,
                                     const int32_t* tok_of_row, const int32_t* tile_expert, const int32_t* n_rows, float* dxrow,
                                     float* dln_w, float* dln_b, void* ws, size_t ws_bytes, int Dm, int E, int row_align,
                                     int64_t max_rows, int dtype, int dxn_dtype, cudaStream_t stream) {
    AB_REQUIRE(ws && ws_bytes >= ab_moe_permute_ln_bwd_workspace_bytes(Dm, row_align, max_rows), "moe_permute_ln_bwd: workspace too small");
    AB_REQUIRE(Dm % 4 == 0, "moe_permute_ln_bwd: hidden size must be a multiple of 4");
    const int sub = work_rows(row_align), ratio = row_align / sub;
    const int ntiles = (int)(max_rows / sub);
    float* part = (float*)ws;
    const int rgrid = rows_grid(max_rows);
    dim3 cgrid(ntiles, (unsigned)ab_ceil_div(Dm, 128));
    const int nit = (int)ab_ceil_div(Dm, 128);
#define AB_LNB_F(TX, TG, NIT)                                                                                                    \
    permute_ln_bwd_fused_kernel<TX, TG, NIT><<<ntiles, 256, 0, stream>>>((const TG*)dxn, (const TX*)x, stats, ln_w, tok_of_row, \
                                                                         tile_expert, n_rows, dxrow, part, Dm, sub, ratio)
#define AB_LNB(TX, TG)                                                                                                           \
    {                                                                                                                            \
        if (nit <= 8) {                                                                                                          \
            if (nit <= 2) AB_LNB_F(TX, TG, 2); else if (nit <= 4) AB_LNB_F(TX, TG, 4);                                           \
            else if (nit <= 6) AB_LNB_F(TX, TG, 6); else AB_LNB_F(TX, TG, 8);                                                    \
        } else {                                                                                                                 \
            permute_ln_bwd_rows_kernel<TX, TG><<<rgrid, 256, 0, stream>>>((const TG*)dxn, (const TX*)x, stats, ln_w, tok_of_row, \
                                                                          tile_expert, n_rows, dxrow, Dm, row_align);            \
            permute_ln_bwd_cols_kernel<TX, TG><<<cgrid, 256, 0, stream>>>((const TG*)dxn, (const TX*)x, stats, tok_of_row,       \
                                                                          tile_expert, n_rows, part, Dm, sub, ratio);            \
        }                                                                                                                        \
    }
    if (dtype == AB_F32 && dxn_dtype == AB_F32) AB_LNB(float, float)
    else if (dtype == AB_F32 && dxn_dtype == AB_BF16) AB_LNB(float, __nv_bfloat16)
    else if (dtype == AB_BF16 && dxn_dtype == AB_BF16) AB_LNB(__nv_bfloat16, __nv_bfloat16)
    else if (dtype == AB_BF16 && dxn_dtype == AB_F32) AB_LNB(__nv_bfloat16, float)
    else AB_REQUIRE(false, "moe_permute_ln_bwd: bad dtypes");
#undef AB_LNB
#undef AB_LNB_F
    AB_LAUNCH_CHECK();
    dim3 grid((unsigned)ab_ceil_div(2 * Dm, 128), E);       // part is [tile][2][Dm]: reduce as 2*Dm columns, split in two
    tile_reduce_kernel<<<grid, 128, 0, stream>>>(part, tile_expert, n_rows, dln_w, dln_b, Dm, 2 * Dm, sub, ratio);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

extern "C" size_t ab_moe_segment_colsum_workspace_bytes(int C, int row_align, int64_t max_rows) {
    return (size_t)ab_round_up((max_rows / work_rows(row_align)) * (int64_t)C * sizeof(float), 256);
}

extern "C" int ab_moe_segment_colsum(const void* a, const int32_t* tile_expert, const int32_t* n_rows, float* out, void* ws,
                                     size_t ws_bytes, int C, int E, int row_align, int64_t max_rows, int dtype,
                                     cudaStream_t stream) {
    AB_REQUIRE(ws && ws_bytes >= ab_moe_segment_colsum_workspace_bytes(C, row_align, max_rows), "moe_segment_colsum: workspace too small");
    const int sub = work_rows(row_align), ratio = row_align / sub;
    const int ntiles = (int)(max_rows / sub);
    float* part = (float*)ws;
    AB_REQUIRE(C % 8 == 0 && ((uintptr_t)a % 16) == 0, "moe_segment_colsum: column count must be a multiple of 8 and the matrix 16-byte aligned");
    dim3 cgrid(ntiles, (unsigned)ab_ceil_div(C, 256));
    if (dtype == AB_F32) tile_colsum_kernel<float><<<cgrid, 256, 0, stream>>>((const float*)a, tile_expert, n_rows, part, C, sub, ratio);
    else tile_colsum_kernel<__nv_bfloat16><<<cgrid, 256, 0, stream>>>((const __nv_bfloat16*)a, tile_expert, n_rows, part, C, sub, ratio);
    AB_LAUNCH_CHECK();
    dim3 grid((unsigned)ab_ceil_div(C, 128), E);
    tile_reduce_kernel<<<grid, 128, 0, stream>>>(part, tile_expert, n_rows, out, nullptr, C, C, sub, ratio);
    AB_LAUNCH_CHECK();
    return AB_OK;
}
