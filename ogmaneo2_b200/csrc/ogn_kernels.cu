// ogn_kernels.cu — hand-written sm_100a kernels of the Sparse Predictive Hierarchy step and their extern "C"
// launchers (include/ogn_kernels_api.h). Compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true
// (-fmad=false: the reference's x86-64 build never contracts a*b+c).
//
// Execution model shared by every kernel: one *lane group* of G = min(32, pow2ceil(Z)) lanes per column, lane = cell
// index (hc for hidden-side kernels, vc for the SparseCoder reconstruction). Each lane accumulates its own cell in
// the reference's sequential slot order (bit-exact sums); the G cells of one (column, slot, active input cell) are
// contiguous in HBM, so a group's load is one coalesced G*4-byte segment. The slot loop is processed in chunks of G:
// every lane first fetches the input CSDR value (or precomputed block offset) of "its" slot of the chunk, the values
// are broadcast with group-masked shuffles, and the chunk's G independent weight loads are issued before the first
// dependent add, which is what keeps enough HBM requests in flight. Column arg-max is a group shuffle reduction with
// the reference's first-max-wins tie-break.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <map>
#include <set>

#include "../../include/ogn_kernels_api.h"
#include "ogn_libm.h"
#include "ogn_stdsort.h"

namespace {

constexpr int kThreads = 256;

// G = lanes per column (pow2ceil(Z), at most 32); P = columns packed per lane group (the layouts' pack factor);
// GP = lanes of one pack group. Must match Geometry::packFactor() on the host.
template <int G> struct Pack {
    static constexpr int P = G >= 8 ? 32 / G : 4;
    static constexpr int GP = G * P;
    static constexpr int B = GP < 16 ? GP : 16; // loads issued per lane before the first dependent add
};

template <int W>
__device__ __forceinline__ unsigned lane_mask() { // mask of the W-lane group this thread belongs to
    if (W == 32) return 0xffffffffu;
    return ((1u << (W & 31)) - 1u) << ((threadIdx.x & 31) & ~(W - 1));
}

__device__ __forceinline__ long long f_block(const ogn_rf &rf, int ox, int q) { // F block of pack (ox, q)
    return (long long)rf.Vz * rf.Hz * rf.pf *
           ((long long)__ldg(rf.px + ox) * rf.syf + (long long)__ldg(rf.nx + ox) * (long long)__ldg(rf.pyf + q));
}
__device__ __forceinline__ long long r_block(const ogn_rf &rf, int ox, int oy) { // R block of hidden column (ox, oy)
    return (long long)rf.Vz * rf.Hz * rf.pr *
           ((long long)__ldg(rf.px + ox) * rf.syr + (long long)__ldg(rf.nx + ox) * (long long)__ldg(rf.pyr + oy));
}
__device__ __forceinline__ long long csr_block(const ogn_rf &rf, int ox, int oy) { // reference CSR order
    return (long long)rf.Vz * rf.Hz *
           ((long long)__ldg(rf.px + ox) * rf.sy + (long long)__ldg(rf.nx + ox) * (long long)__ldg(rf.py + oy));
}

// (value, index) max over a G-lane group; the lowest index wins ties. Values must be NaN-free.
template <int G>
__device__ __forceinline__ void group_argmax(unsigned m, float &v, int &i) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(m, v, o, G);
        int oi = __shfl_xor_sync(m, i, o, G);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

template <bool kReadOnly> __device__ __forceinline__ float load_w(const float *a) { return kReadOnly ? __ldg(a) : *a; }

// a / b the way the reference's x86-64 build rounds AND signals it: a column whose field is empty (radius 0 onto a much
// smaller grid) divides 0 by 0, which SSE answers with the default NaN 0xFFC00000 while the GPU answers 0x7FFFFFFF. Every
// NaN on this path starts at such a division (nothing else is invalid, and x86 propagates the payload it is given), so
// mapping the quotient's NaN to the default NaN keeps even those buffers bit-identical to the reference.
__device__ __forceinline__ float div_ref(float a, float b) {
    const float r = a / b;
    return r != r ? __int_as_float(0xFFC00000) : r;
}

// One pack of P y-adjacent hidden columns seen through one visible layer: the union field the pack group walks and this
// lane's own column's y-range inside it.
struct PackCtx {
    int lx, lyU, nyU, nslotsU; // union field: first visible x / y, y extent, number of union slots
    int dlo, dn;               // this lane's column covers dyU in [dlo, dlo + dn); dn == 0: lane has no column
    long long blk;             // F block offset of the pack
};
__device__ __forceinline__ PackCtx pack_ctx(const ogn_rf &rf, int ox, int q, int p) {
    PackCtx c;
    c.lx = __ldg(rf.lox + ox);
    c.lyU = __ldg(rf.loyf + q);
    c.nyU = __ldg(rf.nyf + q);
    c.nslotsU = __ldg(rf.nx + ox) * c.nyU;
    const int oy = q * rf.pf + p;
    c.dlo = 0; c.dn = 0;
    if (p < rf.pf && oy < rf.Hy) { c.dlo = __ldg(rf.loy + oy) - c.lyU; c.dn = __ldg(rf.ny + oy); }
    c.blk = f_block(rf, ox, q);
    return c;
}

// Σ over this lane's column's receptive field of W_F[...][cs[visible column]][p][hc], in the reference's slot order
// (SparseMatrix::multiplyOHVs, SparseMatrix.cpp:260-276). w = pack block + p*Hz + hc (hc clamped into range).
// The GP lanes of the pack group walk the UNION field together: each lane turns "its" union slot of the current GP slots
// into a line offset (one CSDR read), offsets are broadcast by shuffle and the lines are loaded UNCONDITIONALLY (slot
// index clamped, padding elements are real memory), so a batch of loads is in flight before the first add. Only the
// adds are guarded (slot inside this column's own field), which keeps every column's sum order exact.
// bit j: union slot (chunk start + j) lies inside this lane's own column's field. With P == 1 the union IS the field.
template <int G>
__device__ __forceinline__ unsigned slot_mask(unsigned gm, int dym, const PackCtx &c) {
    constexpr int GP = Pack<G>::GP;
    if (Pack<G>::P == 1) return 0xffffffffu;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < GP; ++j)
        m |= ((unsigned)(__shfl_sync(gm, dym, j, GP) - c.dlo) < (unsigned)c.dn ? 1u : 0u) << j;
    return m;
}

template <int G, bool kReadOnly>
__device__ __forceinline__ float gather_pack(unsigned gm, int gl, const float *w, const int *cs, const PackCtx &c, int Vy, int Vz,
                                             int rowW) {
    constexpr int GP = Pack<G>::GP, B = Pack<G>::B;
    float p = 0.0f;
    for (int ch = 0; ch < c.nslotsU; ch += GP) {
        const int sm = min(ch + gl, c.nslotsU - 1);
        const int dxm = sm / c.nyU, dym = sm - dxm * c.nyU;
        const int cm = __ldg(cs + (c.lx + dxm) * Vy + c.lyU + dym);
        const int om = (sm * Vz + cm) * rowW;
        const int left = c.nslotsU - ch; // valid slots in this chunk (may exceed GP)
        const unsigned vm = slot_mask<G>(gm, dym, c) & (left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
#pragma unroll
        for (int sub = 0; sub < GP; sub += B) {
            if (sub < left) {
                float wv[B];
#pragma unroll
                for (int j = 0; j < B; ++j) wv[j] = load_w<kReadOnly>(w + __shfl_sync(gm, om, sub + j, GP));
#pragma unroll
                for (int j = 0; j < B; ++j)
                    if ((vm >> (sub + j)) & 1u) p += wv[j];
            }
        }
    }
    return p;
}

// W_F[...][cs[...]][p][hc] += delta over this lane's column's field (SparseMatrix::deltaOHVs, SparseMatrix.cpp:488-501).
template <int G>
__device__ __forceinline__ void delta_pack(unsigned gm, int gl, bool act, float *w, const int *cs, float delta, const PackCtx &c, int Vy,
                                           int Vz, int rowW) {
    constexpr int GP = Pack<G>::GP, B = Pack<G>::B;
    for (int ch = 0; ch < c.nslotsU; ch += GP) {
        const int sm = min(ch + gl, c.nslotsU - 1);
        const int dxm = sm / c.nyU, dym = sm - dxm * c.nyU;
        const int cm = __ldg(cs + (c.lx + dxm) * Vy + c.lyU + dym);
        const int om = (sm * Vz + cm) * rowW;
        const int left = c.nslotsU - ch;
        unsigned vm = slot_mask<G>(gm, dym, c) & (left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
        if (!act) vm = 0;
#pragma unroll
        for (int sub = 0; sub < GP; sub += B) {
            if (sub < left) {
                float wv[B];
#pragma unroll
                for (int j = 0; j < B; ++j) wv[j] = w[__shfl_sync(gm, om, sub + j, GP)];
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const int o = __shfl_sync(gm, om, sub + j, GP);
                    if ((vm >> (sub + j)) & 1u) w[o] = wv[j] + delta;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2: SparseCoder::forward (SparseCoder.cpp:13-43). One pack group per hidden-y pack.
template <int G>
__global__ void __launch_bounds__(kThreads)
k_sc_forward(const __grid_constant__ ogn_fwd_args A, int Hx, int Hy, int Hz, int x0, int npacks, int packsPerRow,
             int *__restrict__ out0, int *__restrict__ out1) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP;
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return; // whole pack groups leave together
    const unsigned gm = lane_mask<GP>(), sm = lane_mask<G>();
    const int member = blockIdx.y;
    const int ox = x0 + pack / packsPerRow, q = pack % packsPerRow;
    const int p = gl / G, hl = gl % G;
    const int oy = q * P + p;
    const bool colValid = oy < Hy;

    int bestIdx = 0;
    float bestVal = -999999.0f;
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + hl;
        const bool act = hc < Hz && colValid;
        const int hcl = hc < Hz ? hc : Hz - 1;
        float sum = 0.0f;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_fwd_vl &vl = A.vl[v];
            const ogn_rf &rf = A.geo[vl.geo];
            const PackCtx c = pack_ctx(rf, ox, q, p);
            const float *w = vl.wf + (c.blk + (long long)member * vl.mstride) + p * Hz + hcl;
            const int *cs = vl.cs + (long long)member * rf.Vx * rf.Vy;
            sum += gather_pack<G, true>(gm, gl, w, cs, c, rf.Vy, rf.Vz, P * Hz);
        }
        // sequential rule: `if (sum > max)` from (idx 0, -999999); NaN never wins
        float v = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
        int i = act ? hc : 0x7fffffff;
        group_argmax<G>(sm, v, i);
        if (v > bestVal) { bestVal = v; bestIdx = i; }
    }
    if (hl == 0 && colValid) {
        const long long o = (long long)member * Hx * Hy + ox * Hy + oy;
        out0[o] = bestIdx;
        if (out1) out1[o] = bestIdx;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: SparseCoder::learn (SparseCoder.cpp:45-83). One pack group per group of pr y-adjacent visible columns;
// lane = (visible column of the pack, visible cell vc).
struct CoverItem { long long offR; int mask; }; // line offset in R, bit pv set: hidden column covers visible column pv

template <int G>
__device__ __forceinline__ CoverItem cover_item(const ogn_rf &rf, int k, int ncyU, int hx0, int hy0U, int vx, int vyBase,
                                                const int *hcs) {
    constexpr int P = Pack<G>::P;
    CoverItem it;
    const int ox = hx0 + k / ncyU, oy = hy0U + k % ncyU;
    const int hc = __ldg(hcs + ox * rf.Hy + oy);
    const int ylo = __ldg(rf.loy + oy), yn = __ldg(rf.ny + oy);
    const int dx = vx - __ldg(rf.lox + ox);
    const int gi = vyBase / P - ylo / P; // visible-y group index inside this hidden column's field
    it.offR = r_block(rf, ox, oy) + (((long long)dx * __ldg(rf.ngr + oy) + gi) * rf.Hz + hc) * (P * rf.Vz);
    it.mask = 0;
#pragma unroll
    for (int pv = 0; pv < P; ++pv)
        if ((unsigned)(vyBase + pv - ylo) < (unsigned)yn && vyBase + pv < rf.Vy) it.mask |= 1 << pv;
    return it;
}

template <int G>
__global__ void __launch_bounds__(kThreads)
k_sc_learn(const __grid_constant__ ogn_learn_args A, const int *__restrict__ hcsAll, const int *__restrict__ hcsPrevAll, int vx0,
           int npacks, int packsPerRow, float alpha, unsigned long long *updCount) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP, B = Pack<G>::B;
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return;
    const unsigned gm = lane_mask<GP>(), sm = lane_mask<G>();
    const int member = blockIdx.y;
    const ogn_learn_vl &vl = A.vl[blockIdx.z];
    const ogn_rf &rf = A.geo;
    const int Vz = rf.Vz;
    const int vx = vx0 + pack / packsPerRow, vyBase = (pack % packsPerRow) * P;
    const int pv = gl / G, vl_ = gl % G;
    const int vy = vyBase + pv;
    const bool colValid = vy < rf.Vy;
    const int vyc = colValid ? vy : rf.Vy - 1;
    const long long vcol = (long long)member * rf.Vx * rf.Vy + vx * rf.Vy + vyc;
    const int target = __ldg(vl.cs + vcol);
    const int *hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
    const int *hcsPrev = hcsPrevAll + (long long)member * rf.Hx * rf.Hy;
    float *wr = vl.wr + (long long)member * vl.r_mstride;
    float *wf = vl.wf + (long long)member * vl.f_mstride;
    float *recon = vl.recon + vcol * Vz;

    // cover sets: own (for the count and the update) and the pack's union along y (walked together)
    const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx);
    const int hy0 = __ldg(rf.rloy + vyc), hy1 = __ldg(rf.rhiy + vyc);
    const int ncx = max(0, hx1 - hx0 + 1), ncy = max(0, hy1 - hy0 + 1);
    const int ncov = ncx * ncy; // == countT(visibleIndex) / hiddenSize.z
    int hy0U = rf.Hy, hy1U = -1;
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int y = vyBase + k;
        if (y < rf.Vy) {
            const int a = __ldg(rf.rloy + y), b = __ldg(rf.rhiy + y);
            if (b >= a) { hy0U = min(hy0U, a); hy1U = max(hy1U, b); }
        }
    }
    const int ncyU = max(0, hy1U - hy0U + 1), ncovU = ncx * ncyU;

    // reconstruction + arg-max with the reference's `sum > max || maxIndex == -1` rule
    int bestIdx = -1;
    float bestVal = -999999.0f;
    float myRecon = 0.0f; // recon of vc = lane's cell when Vz <= G
    for (int vc0 = 0; vc0 < Vz; vc0 += G) {
        const int vc = vc0 + vl_;
        const bool act = vc < Vz && colValid;
        const int vcl = vc < Vz ? vc : Vz - 1;
        float sum = 0.0f;
        for (int ch = 0; ch < ncovU; ch += GP) {
            const CoverItem mine = cover_item<G>(rf, min(ch + gl, ncovU - 1), ncyU, hx0, hy0U, vx, vyBase, hcs);
            const int left = ncovU - ch;
            // bit j: hidden column (chunk start + j) covers this lane's visible column
            unsigned vm = 0xffffffffu;
            if (P != 1) {
                vm = 0;
#pragma unroll
                for (int j = 0; j < GP; ++j) vm |= (unsigned)((__shfl_sync(gm, mine.mask, j, GP) >> pv) & 1) << j;
            }
            vm &= left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
#pragma unroll
            for (int sub = 0; sub < GP; sub += B) {
                if (sub < left) {
                    float wv[B];
#pragma unroll
                    for (int j = 0; j < B; ++j) wv[j] = wr[__shfl_sync(gm, mine.offR, sub + j, GP) + pv * Vz + vcl];
#pragma unroll
                    for (int j = 0; j < B; ++j)
                        if ((vm >> (sub + j)) & 1u) sum += wv[j];
                }
            }
        }
        sum = div_ref(sum, (float)ncov);
        if (act) recon[vc] = sum;
        if (vc0 == 0) myRecon = sum;
        // chunk arg-max: element 0 of the whole column is always taken first, NaN elsewhere never wins
        float val;
        if (!act) val = -INFINITY;
        else if (vc == 0) val = (sum != sum) ? INFINITY : sum;
        else val = (sum != sum) ? -INFINITY : sum;
        int idx = act ? vc : 0x7fffffff;
        group_argmax<G>(sm, val, idx);
        if (bestIdx == -1 || val > bestVal) { bestVal = val; bestIdx = idx; }
    }
    if (!colValid || bestIdx == target) return; // from here on the G lanes of one visible column work on their own

    // delta rule on the weights of hidden columns whose state changed (deltaChangedOHVsT, SparseMatrix.cpp:572-590)
    int nupd = 0;
    for (int vc0 = 0; vc0 < Vz; vc0 += G) {
        const int vc = vc0 + vl_;
        const bool act = vc < Vz;
        float r = myRecon;
        if (vc0 != 0 && act) r = recon[vc];
        const float delta = alpha * ((vc == target ? 1.0f : 0.0f) - ogn_expf(r));
        for (int ch = 0; ch < ncov; ch += G) {
            // lane's item: one hidden column of this visible column's own cover set
            long long offR = 0, offF = 0;
            int fl = 0;
            if (ch + vl_ < ncov) {
                const int k = ch + vl_;
                const int ox = hx0 + k / ncy, oy = hy0 + k % ncy;
                const int h = ox * rf.Hy + oy;
                const int hc = __ldg(hcs + h);
                const int ylo = __ldg(rf.loy + oy);
                const int dx = vx - __ldg(rf.lox + ox);
                offR = r_block(rf, ox, oy) +
                       ((((long long)dx * __ldg(rf.ngr + oy) + (vy / P - ylo / P)) * rf.Hz + hc) * P + (vy % P)) * Vz;
                const int qh = oy / rf.pf, ph = oy % rf.pf;
                offF = f_block(rf, ox, qh) +
                       (((long long)dx * __ldg(rf.nyf + qh) + (vy - __ldg(rf.loyf + qh))) * Vz * rf.pf + ph) * rf.Hz + hc;
                fl = (hc != __ldg(hcsPrev + h) ? 1 : 0) | ((ox >= A.fx0 && ox < A.fx1) ? 2 : 0);
            }
#pragma unroll 4
            for (int j = 0; j < G; ++j) {
                const long long oR = __shfl_sync(sm, offR, j, G);
                const long long oF = __shfl_sync(sm, offF, j, G);
                const int f = __shfl_sync(sm, fl, j, G);
                if (f & 1) {
                    if (act) {
                        const float nw = wr[oR + vc] + delta;
                        wr[oR + vc] = nw;
                        if (f & 2) wf[oF + (long long)vc * rf.pf * rf.Hz] = nw;
                    }
                    if (vc0 == 0) nupd++;
                }
            }
        }
    }
    if (updCount && vl_ == 0 && nupd) atomicAdd(updCount, (unsigned long long)nupd);
}

// ---------------------------------------------------------------------------------------------------------------
// K6+K7+K8: Predictor::learn (Predictor.cpp:49-70), Predictor::forward (:13-47), inputCs -> inputCsPrev (:118-126)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_pred_step(const __grid_constant__ ogn_pred_args A, int Hx, int Hy, int Hz, int x0, int npacks, int packsPerRow, int learn, int activate,
            const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP;
    const int member = blockIdx.y;
    // K8: every thread of the grid helps copying the inputs into the *other* inputCsPrev buffer (never read here)
    if (activate == 1) {
        const int tid = blockIdx.x * kThreads + threadIdx.x, nth = gridDim.x * kThreads;
        for (int v = 0; v < A.nvl; ++v) {
            if (!A.vl[v].cs_prev_out) continue;
            const int n = A.geo[A.vl[v].geo].Vx * A.geo[A.vl[v].geo].Vy;
            const long long mo = (long long)member * n;
            for (int i = tid; i < n; i += nth) A.vl[v].cs_prev_out[mo + i] = __ldg(A.vl[v].cs + mo + i);
        }
    }
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return;
    const unsigned gm = lane_mask<GP>(), sm = lane_mask<G>();
    const int ox = x0 + pack / packsPerRow, q = pack % packsPerRow;
    const int p = gl / G, hl = gl % G;
    const int oy = q * P + p;
    const bool colValid = oy < Hy;
    const long long col = (long long)member * Hx * Hy + ox * Hy + (colValid ? oy : Hy - 1);

    if (learn) {
        const int t = __ldg(target + col);
        for (int hc0 = 0; hc0 < Hz; hc0 += G) {
            const int hc = hc0 + hl;
            const bool act = hc < Hz && colValid;
            const int hcl = hc < Hz ? hc : Hz - 1;
            float delta = 0.0f;
            if (act) delta = alpha * ((hc == t ? 1.0f : -1.0f) - ogn_tanhf(acts[col * Hz + hc]));
            for (int v = 0; v < A.nvl; ++v) {
                const ogn_pred_vl &vl = A.vl[v];
                const ogn_rf &rf = A.geo[vl.geo];
                const PackCtx c = pack_ctx(rf, ox, q, p);
                float *w = vl.wf + (c.blk + (long long)member * vl.mstride) + p * Hz + hcl;
                const int *cs = vl.cs_prev + (long long)member * rf.Vx * rf.Vy;
                delta_pack<G>(gm, gl, act, w, cs, delta, c, rf.Vy, rf.Vz, P * Hz);
            }
        }
    }
    if (!activate) return;

    int bestIdx = 0;
    float bestVal = -999999.0f;
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + hl;
        const bool act = hc < Hz && colValid;
        const int hcl = hc < Hz ? hc : Hz - 1;
        float sum = 0.0f;
        int count = 0;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_pred_vl &vl = A.vl[v];
            const ogn_rf &rf = A.geo[vl.geo];
            const PackCtx c = pack_ctx(rf, ox, q, p);
            const float *w = vl.wf + (c.blk + (long long)member * vl.mstride) + p * Hz + hcl;
            const int *cs = vl.cs + (long long)member * rf.Vx * rf.Vy;
            sum += gather_pack<G, false>(gm, gl, w, cs, c, rf.Vy, rf.Vz, P * Hz);
            count += c.dn * (c.nslotsU / max(c.nyU, 1)); // own field: nx * ny
        }
        sum = div_ref(sum, (float)count);
        if (act) acts[col * Hz + hc] = sum;
        float v = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
        int i = act ? hc : 0x7fffffff;
        group_argmax<G>(sm, v, i);
        if (v > bestVal) { bestVal = v; bestIdx = i; }
    }
    if (hl == 0 && colValid) hiddenCs[col] = bestIdx;
}

// ---------------------------------------------------------------------------------------------------------------
// Actor. One pack group per hidden-y pack of action columns (template G from the action Z). The value matrix has
// Hz = 1, i.e. its own geometry with pack factor 4: lane 0 of every column gathers the value row with the generic
// single-lane routines below (action layers are tiny; SURVEY.md: "negligible bytes, correctness-critical RNG").
// Value row of one action column (the value matrix has Hz = 1): Σ over the column's field of W_v[slot][cs[slot]] in slot
// order (Actor.cpp:27-33 / :105-111). The G lanes of the column's group cooperate: lane s % G loads slot s (its CSDR
// read and its weight, 8 independent chains in flight per lane), the ordered sum is rebuilt with shuffles, so every lane
// returns the same bits the sequential loop produces.
template <int G>
__device__ __forceinline__ float value_row_sum(unsigned sm, int hl, bool valid, const ogn_rf &rf, const float *wv, const int *cs, int ox,
                                               int oy, int *countOut) {
    const int q = oy / rf.pf, p = oy % rf.pf;
    const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
    const int lyU = __ldg(rf.loyf + q), nyU = __ldg(rf.nyf + q);
    const float *w = wv + f_block(rf, ox, q) + p;
    const int n = valid ? nxv * nyv : 0;
    float sum = 0.0f;
    for (int s0 = 0; s0 < n; s0 += G * 8) {
        float w8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int s = s0 + k * G + hl;
            w8[k] = 0.0f;
            if (s < n) {
                const int dx = s / nyv, dy = s - dx * nyv;
                const int c = __ldg(cs + (lx + dx) * rf.Vy + ly + dy);
                w8[k] = w[((long long)(dx * nyU + (ly + dy - lyU)) * rf.Vz + c) * rf.pf];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const float v = __shfl_sync(sm, w8[k], j, G);
                if (s0 + k * G + j < n) sum += v;
            }
    }
    *countOut = nxv * nyv;
    return sum;
}
// W_v[slot][cs[slot]] += delta over the column's field (Actor.cpp:117-118); lane s % G owns slot s, as above
template <int G>
__device__ __forceinline__ void value_row_delta(int hl, bool valid, const ogn_rf &rf, float *wv, const int *cs, int ox, int oy, float delta) {
    const int q = oy / rf.pf, p = oy % rf.pf;
    const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
    const int lyU = __ldg(rf.loyf + q), nyU = __ldg(rf.nyf + q);
    float *w = wv + f_block(rf, ox, q) + p;
    const int n = valid ? nxv * nyv : 0;
    for (int s = hl; s < n; s += G) {
        const int dx = s / nyv, dy = s - dx * nyv;
        const int c = __ldg(cs + (lx + dx) * rf.Vy + ly + dy);
        w[((long long)(dx * nyU + (ly + dy - lyU)) * rf.Vz + c) * rf.pf] += delta;
    }
}

// value of the column (identical on every lane of its group) and the total slot count; cs0 != nullptr overrides the
// visible layers' CSDR pointers (history samples)
template <int G>
__device__ __forceinline__ float actor_value(unsigned sm, int hl, bool colValid, const ogn_actor_args &A, const int *const *csv, int ox,
                                             int oy, int *countOut) {
    float value = 0.0f;
    int count = 0;
    for (int v = 0; v < A.nvl; ++v) {
        int n;
        value += value_row_sum<G>(sm, hl, colValid, A.geo[A.vl[v].geo_v], A.vl[v].wv, csv[v], ox, oy, &n);
        count += n;
    }
    *countOut = count;
    return value;
}

// raw action activations (Σ_v gather) / count into acts[] (column base)
// Returns the lane's own activation of the first chunk (hc = hl), which is all of them when Hz <= G; acts[] is written
// only when Hz > G (the softmax then goes through that scratch).
template <int G>
__device__ __forceinline__ float actor_action_acts(unsigned gm, unsigned sm, int gl, int p, int hl, bool colValid, const ogn_actor_args &A,
                                                   const int *const *csv, int ox, int q, int Hz, int count, float *acts) {
    constexpr int P = Pack<G>::P;
    float mine = 0.0f;
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + hl;
        const bool act = hc < Hz && colValid;
        const int hcl = hc < Hz ? hc : Hz - 1;
        float sum = 0.0f;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo_a];
            const PackCtx c = pack_ctx(rf, ox, q, p);
            const float *w = A.vl[v].wa + c.blk + p * Hz + hcl;
            sum += gather_pack<G, false>(gm, gl, w, csv[v], c, rf.Vy, rf.Vz, P * Hz);
        }
        sum = div_ref(sum, (float)count);
        if (hc0 == 0) mine = sum;
        if (act && Hz > G) acts[hc] = sum;
    }
    __syncwarp(sm);
    return mine;
}

// The same softmax when the column's Hz cells sit one per lane (Hz <= G): the max and the ascending-order sum are rebuilt
// with shuffles, nothing goes through memory. e = this lane's numerator expf(act - max) (0 for lanes past Hz).
template <int G>
__device__ __forceinline__ float actor_softmax_lanes(unsigned sm, int hl, int Hz, float act, float &e) {
    float m = -999999.0f;
    for (int hc = 0; hc < Hz; ++hc) m = fmaxf(m, __shfl_sync(sm, act, hc, G));
    e = hl < Hz ? ogn_expf(act - m) : 0.0f;
    float total = 0.0f;
    for (int hc = 0; hc < Hz; ++hc) total += __shfl_sync(sm, e, hc, G);
    return total;
}

// softmax numerators in place (Actor.cpp:57-69 / :159-171): acts[hc] = expf(acts[hc] - max); returns Σ in ascending hc order.
template <int G>
__device__ __forceinline__ float actor_softmax(unsigned sm, int hl, int Hz, float *acts) {
    float m = -999999.0f;
    for (int hc = 0; hc < Hz; ++hc) m = fmaxf(m, __ldcg(acts + hc)); // every lane scans the same Hz values (L2 reads)
    __syncwarp(sm);
    for (int hc = hl; hc < Hz; hc += G) acts[hc] = ogn_expf(__ldcg(acts + hc) - m);
    __syncwarp(sm);
    float total = 0.0f;
    for (int hc = 0; hc < Hz; ++hc) total += __ldcg(acts + hc);
    return total;
}

// K9: Actor::forward (Actor.cpp:13-90)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_actor_forward(const __grid_constant__ ogn_actor_args A, int Hx, int Hy, int Hz, int npacks, int packsPerRow,
                const float *__restrict__ u01, float *hiddenValues, float *hiddenActs, int *hiddenCs) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP;
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return;
    const unsigned gm = lane_mask<GP>(), sm = lane_mask<G>();
    const int ox = pack / packsPerRow, q = pack % packsPerRow;
    const int p = gl / G, hl = gl % G;
    const int oy = q * P + p;
    const bool colValid = oy < Hy;
    const int oyc = colValid ? oy : Hy - 1;
    const int col = ox * Hy + oyc;
    const int *csv[OGN_MAX_PRED_VL];
    for (int v = 0; v < A.nvl; ++v) csv[v] = A.vl[v].cs;
    int count;
    const float value = actor_value<G>(sm, hl, colValid, A, csv, ox, oyc, &count);
    if (hl == 0 && colValid) hiddenValues[col] = div_ref(value, (float)count);
    float *acts = hiddenActs + (long long)col * Hz;
    const float mine = actor_action_acts<G>(gm, sm, gl, p, hl, colValid, A, csv, ox, q, Hz, count, acts);
    if (!colValid) return;
    if (Hz <= G) {
        float e;
        const float total = actor_softmax_lanes<G>(sm, hl, Hz, mine, e);
        if (hl < Hz) acts[hl] = e; // hiddenActivations keeps the numerators (Actor.cpp:66)
        // std::uniform_real_distribution<float>(0, total): canonical * (b - a) + a   (libstdc++ bits/random.h)
        const float cusp = u01[col] * (total - 0.0f) + 0.0f;
        int select = 0;
        float sumSoFar = 0.0f;
        bool found = false;
        for (int hc = 0; hc < Hz; ++hc) { // every lane replays the scan (the shuffles need all of them)
            sumSoFar += __shfl_sync(sm, e, hc, G);
            if (!found && sumSoFar >= cusp) { select = hc; found = true; }
        }
        if (hl == 0) hiddenCs[col] = select;
        return;
    }
    const float total = actor_softmax<G>(sm, hl, Hz, acts);
    if (hl == 0) {
        const float cusp = u01[col] * (total - 0.0f) + 0.0f;
        int select = 0;
        float sumSoFar = 0.0f;
        for (int hc = 0; hc < Hz; ++hc) {
            sumSoFar += __ldcg(acts + hc);
            if (sumSoFar >= cusp) { select = hc; break; }
        }
        hiddenCs[col] = select;
    }
}

// K11: Actor::learn (Actor.cpp:92-185) on B.n history samples, one after the other, in ONE launch: every action column
// owns its value row and its action rows, so a column's lanes can walk the samples on their own (the reference launches
// one kernel per sample; the result is the same because nothing crosses columns).
template <int G>
__global__ void __launch_bounds__(kThreads)
k_actor_learn(const __grid_constant__ ogn_actor_args A, const __grid_constant__ ogn_actor_batch B, int Hx, int Hy, int Hz, int npacks,
              int packsPerRow, const float *__restrict__ hiddenValues, float alpha, float beta, int mimic, float *hiddenActs) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP;
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return;
    const unsigned gm = lane_mask<GP>(), sm = lane_mask<G>();
    const int ox = pack / packsPerRow, q = pack % packsPerRow;
    const int p = gl / G, hl = gl % G;
    const int oy = q * P + p;
    const bool colValid = oy < Hy;
    const int oyc = colValid ? oy : Hy - 1;
    const int col = ox * Hy + oyc;
    float *acts = hiddenActs + (long long)col * Hz;

    for (int it = 0; it < B.n; ++it) {
        const ogn_actor_sample &S = B.it[it];
        const float newValue = S.q + S.g * hiddenValues[col];
        int count;
        float value = actor_value<G>(sm, hl, colValid, A, S.cs, ox, oyc, &count);
        value = div_ref(value, (float)count);
        const float tdErrorValue = newValue - value;
        const float deltaValue = alpha * tdErrorValue;
        for (int v = 0; v < A.nvl; ++v) value_row_delta<G>(hl, colValid, A.geo[A.vl[v].geo_v], A.vl[v].wv, S.cs[v], ox, oyc, deltaValue);
        const float tdErrorAction = newValue - S.values_prev[col];
        const int target = S.target_prev[col];
        const float mine = actor_action_acts<G>(gm, sm, gl, p, hl, colValid, A, S.cs, ox, q, Hz, count, acts);
        float total = 1.0f, e = 0.0f;
        const bool lanes = Hz <= G; // one cell per lane: softmax through shuffles, else through the acts[] scratch
        if (lanes) {
            total = actor_softmax_lanes<G>(sm, hl, Hz, mine, e);
            if (hl < Hz && colValid) acts[hl] = e;
        } else if (colValid)
            total = actor_softmax<G>(sm, hl, Hz, acts);
        for (int hc0 = 0; hc0 < Hz; hc0 += G) {
            const int hc = hc0 + hl;
            const bool act = hc < Hz && colValid;
            const int hcl = hc < Hz ? hc : Hz - 1;
            float delta = 0.0f;
            if (act)
                delta = (mimic ? beta : (tdErrorAction > 0.0f ? beta : -beta)) *
                        ((hc == target ? 1.0f : 0.0f) - (lanes ? e : __ldcg(acts + hc)) / fmaxf(0.0001f, total));
            for (int v = 0; v < A.nvl; ++v) {
                const ogn_rf &rf = A.geo[A.vl[v].geo_a];
                const PackCtx c = pack_ctx(rf, ox, q, p);
                float *w = A.vl[v].wa + c.blk + p * Hz + hcl;
                delta_pack<G>(gm, gl, act, w, S.cs[v], delta, c, rf.Vy, rf.Vz, P * Hz);
            }
        }
        __syncwarp(gm); // this sample's weight and scratch writes are ordered before the next sample's reads
    }
}

// K11, latency version: one BLOCK per action column, one WARP per weight row of the column (warp 0: the value row, warp
// 1 + hc: action row hc). A row's gather touches one weight per receptive-field slot; the lanes of the row's warp fetch
// them all at once (one round trip through L2 instead of one per slot), and the sum is rebuilt in slot order, visible
// layer by visible layer, with shuffles, so it is the float the sequential loop produces (Actor.cpp:105-111, :141-151).
// The history samples of a step are still applied one after the other (a sample's gathers must see the previous sample's
// deltas), but a sample now costs a few dependent L2 round trips instead of ~6.6 k dependent instructions.
struct ActorRowGeo { // own field of the column inside one visible layer, seen through one matrix' packed layout
    int lx, ly, nx, ny, lyU, nyU, Vy, Vz;
    long long stride; // floats between consecutive (union slot, visible cell) entries of this row
};
__device__ __forceinline__ ActorRowGeo actor_row_geo(const ogn_rf &rf, int ox, int oy, long long rowStride) {
    ActorRowGeo g;
    const int q = oy / rf.pf;
    g.lx = __ldg(rf.lox + ox); g.nx = __ldg(rf.nx + ox); g.ly = __ldg(rf.loy + oy); g.ny = __ldg(rf.ny + oy);
    g.lyU = __ldg(rf.loyf + q); g.nyU = __ldg(rf.nyf + q); g.Vy = rf.Vy; g.Vz = rf.Vz; g.stride = rowStride;
    return g;
}
// ordered sum over the field of W[row][slot][cs[slot]] (multiplyOHVs); identical on every lane
__device__ __forceinline__ float actor_row_gather(int lane, const float *w, const int *cs, const ActorRowGeo &g) {
    const int n = g.nx * g.ny;
    float sum = 0.0f;
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + lane;
        float v = 0.0f;
        if (s < n) {
            const int dx = s / g.ny, dy = s - dx * g.ny;
            const int c = cs[(g.lx + dx) * g.Vy + g.ly + dy];
            v = w[((long long)(dx * g.nyU + (g.ly + dy - g.lyU)) * g.Vz + c) * g.stride];
        }
        const int m = min(32, n - s0);
        for (int j = 0; j < m; ++j) sum += __shfl_sync(0xffffffffu, v, j);
    }
    return sum;
}
// W[row][slot][cs[slot]] += delta over the field (deltaOHVs): every slot is a different element, one lane each
__device__ __forceinline__ void actor_row_delta(int lane, float *w, const int *cs, const ActorRowGeo &g, float delta) {
    const int n = g.nx * g.ny;
    for (int s = lane; s < n; s += 32) {
        const int dx = s / g.ny, dy = s - dx * g.ny;
        const int c = cs[(g.lx + dx) * g.Vy + g.ly + dy];
        w[((long long)(dx * g.nyU + (g.ly + dy - g.lyU)) * g.Vz + c) * g.stride] += delta;
    }
}

__global__ void __launch_bounds__(1024)
k_actor_learn_rows(const __grid_constant__ ogn_actor_args A, const __grid_constant__ ogn_actor_batch B, int Hx, int Hy, int Hz,
                   const float *__restrict__ hiddenValues, float alpha, float beta, int mimic, float *hiddenActs) {
    __shared__ float sActs[32];
    __shared__ float sNewValue;
    const int col = blockIdx.x, ox = col / Hy, oy = col - ox * Hy;
    const int row = threadIdx.x >> 5, lane = threadIdx.x & 31; // row 0: value, row 1 + hc: action cell hc
    const bool isValue = row == 0;
    const int hc = row - 1;
    for (int it = 0; it < B.n; ++it) {
        const ogn_actor_sample &S = B.it[it];
        float sum = 0.0f;
        int count = 0;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[isValue ? A.vl[v].geo_v : A.vl[v].geo_a];
            const int q = oy / rf.pf, p = oy % rf.pf;
            const ActorRowGeo g = actor_row_geo(rf, ox, oy, isValue ? (long long)rf.pf : (long long)rf.pf * Hz);
            const float *w = isValue ? A.vl[v].wv + f_block(rf, ox, q) + p : A.vl[v].wa + f_block(rf, ox, q) + (long long)p * Hz + hc;
            sum += actor_row_gather(lane, w, S.cs[v], g);
            count += g.nx * g.ny;
        }
        sum = div_ref(sum, (float)count);
        if (isValue) {
            const float newValue = S.q + S.g * hiddenValues[col];
            if (lane == 0) sNewValue = newValue;
            const float deltaValue = alpha * (newValue - sum);
            for (int v = 0; v < A.nvl; ++v) {
                const ogn_rf &rf = A.geo[A.vl[v].geo_v];
                const int q = oy / rf.pf, p = oy % rf.pf;
                actor_row_delta(lane, A.vl[v].wv + f_block(rf, ox, q) + p, S.cs[v], actor_row_geo(rf, ox, oy, (long long)rf.pf), deltaValue);
            }
        } else if (lane == 0)
            sActs[hc] = sum;
        __syncthreads();
        if (!isValue) {
            // softmax over the column's Hz activations in the reference's order (Actor.cpp:153-171): every action warp
            // rebuilds max and total itself (lane h holds cell h), so no second block barrier is needed
            float m = -999999.0f;
            for (int h = 0; h < Hz; ++h) m = fmaxf(m, sActs[h]);
            const float e = lane < Hz ? ogn_expf(sActs[lane] - m) : 0.0f;
            float total = 0.0f;
            for (int h = 0; h < Hz; ++h) total += __shfl_sync(0xffffffffu, e, h);
            const float mine = __shfl_sync(0xffffffffu, e, hc);
            const float tdErrorAction = sNewValue - S.values_prev[col];
            const int target = S.target_prev[col];
            const float deltaAction = (mimic ? beta : (tdErrorAction > 0.0f ? beta : -beta)) * ((hc == target ? 1.0f : 0.0f) - mine / fmaxf(0.0001f, total));
            for (int v = 0; v < A.nvl; ++v) {
                const ogn_rf &rf = A.geo[A.vl[v].geo_a];
                const int q = oy / rf.pf, p = oy % rf.pf;
                actor_row_delta(lane, A.vl[v].wa + f_block(rf, ox, q) + (long long)p * Hz + hc, S.cs[v],
                                actor_row_geo(rf, ox, oy, (long long)rf.pf * Hz), deltaAction);
            }
            if (lane == 0) hiddenActs[(long long)col * Hz + hc] = mine; // hiddenActivations keeps the numerators (Actor.cpp:164)
        }
        __syncthreads(); // sActs / sNewValue are rewritten by the next sample; a row's weights stay with its own warp
    }
}

// K9 in the same mapping (Actor::forward, Actor.cpp:13-90): value row and action rows gathered by one warp each, softmax and
// the sampling scan on the action warps (every one of them rebuilds the column's numerators; warp 1 writes the result).
__global__ void __launch_bounds__(1024)
k_actor_forward_rows(const __grid_constant__ ogn_actor_args A, int Hx, int Hy, int Hz, const float *__restrict__ u01, float *hiddenValues,
                     float *hiddenActs, int *hiddenCs) {
    __shared__ float sActs[32];
    const int col = blockIdx.x, ox = col / Hy, oy = col - ox * Hy;
    const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool isValue = row == 0;
    const int hc = row - 1;
    float sum = 0.0f;
    int count = 0;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[isValue ? A.vl[v].geo_v : A.vl[v].geo_a];
        const int q = oy / rf.pf, p = oy % rf.pf;
        const ActorRowGeo g = actor_row_geo(rf, ox, oy, isValue ? (long long)rf.pf : (long long)rf.pf * Hz);
        const float *w = isValue ? A.vl[v].wv + f_block(rf, ox, q) + p : A.vl[v].wa + f_block(rf, ox, q) + (long long)p * Hz + hc;
        sum += actor_row_gather(lane, w, A.vl[v].cs, g);
        count += g.nx * g.ny;
    }
    sum = div_ref(sum, (float)count);
    if (isValue) { if (lane == 0) hiddenValues[col] = sum; }
    else if (lane == 0) sActs[hc] = sum;
    __syncthreads();
    if (isValue) return;
    float m = -999999.0f;
    for (int h = 0; h < Hz; ++h) m = fmaxf(m, sActs[h]);
    const float e = lane < Hz ? ogn_expf(sActs[lane] - m) : 0.0f;
    float total = 0.0f;
    for (int h = 0; h < Hz; ++h) total += __shfl_sync(0xffffffffu, e, h);
    if (hc != 0) return; // one writer per column
    if (lane < Hz) hiddenActs[(long long)col * Hz + lane] = e; // hiddenActivations keeps the numerators (Actor.cpp:66)
    // std::uniform_real_distribution<float>(0, total): canonical * (b - a) + a   (libstdc++ bits/random.h)
    const float cusp = u01[col] * (total - 0.0f) + 0.0f;
    int select = 0;
    float sumSoFar = 0.0f;
    bool found = false;
    for (int h = 0; h < Hz; ++h) {
        sumSoFar += __shfl_sync(0xffffffffu, e, h);
        if (!found && sumSoFar >= cusp) { select = h; found = true; }
    }
    if (lane == 0) hiddenCs[col] = select;
}

// ---------------------------------------------------------------------------------------------------------------
// ImageEncoder (ImageEncoder.cpp:19-94). One pack group per hidden-y pack; lane = (column p of the pack, hidden cell). The
// F block of a pack is read front to back: rows (union slot, visible cell) are consecutive, each holding the pack's
// P * Hz cells, so a group's loads are contiguous 128-byte segments. Every cell keeps its own running sum in the
// reference's order (slots x-major, then y, then visible cell; visible layer after visible layer).
template <int G>
__device__ __forceinline__ float ienc_distance2(const float *w, const float *in, const PackCtx &c, const ogn_rf &rf, int rowW, bool act) {
    float d2 = 0.0f;
    const int nxv = c.nyU > 0 ? c.nslotsU / c.nyU : 0;
    for (int dx = 0; dx < nxv; ++dx)
        for (int dy = 0; dy < c.nyU; ++dy) {
            const bool own = act && (unsigned)(dy - c.dlo) < (unsigned)c.dn;
            const float *wr = w + (long long)((dx * c.nyU + dy) * rf.Vz) * rowW;
            const float *ir = in + (long long)rf.Vz * ((c.lyU + dy) + (long long)rf.Vy * (c.lx + dx));
            for (int vz0 = 0; vz0 < rf.Vz; vz0 += 8) {
                float wv[8], iv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int vz = min(vz0 + k, rf.Vz - 1);
                    wv[k] = wr[(long long)vz * rowW];
                    iv[k] = own ? __ldg(ir + vz) : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (own && vz0 + k < rf.Vz) { const float d = iv[k] - wv[k]; d2 += d * d; }
            }
        }
    return d2;
}
template <int G>
__device__ __forceinline__ void ienc_hebb(float *w, const float *in, const PackCtx &c, const ogn_rf &rf, int rowW, bool act, float strength) {
    const int nxv = c.nyU > 0 ? c.nslotsU / c.nyU : 0;
    for (int dx = 0; dx < nxv; ++dx)
        for (int dy = 0; dy < c.nyU; ++dy) {
            if (!(act && (unsigned)(dy - c.dlo) < (unsigned)c.dn)) continue;
            float *wr = w + (long long)((dx * c.nyU + dy) * rf.Vz) * rowW;
            const float *ir = in + (long long)rf.Vz * ((c.lyU + dy) + (long long)rf.Vy * (c.lx + dx));
            for (int vz = 0; vz < rf.Vz; ++vz) {
                const float old = wr[(long long)vz * rowW];
                wr[(long long)vz * rowW] = old + strength * (__ldg(ir + vz) - old);
            }
        }
}

// rank of this cell among the column's cells, activations descending. The reference ranks with std::sort and a
// comparator on the value only (ImageEncoder.cpp:15-17, :56): without equal activations every sort gives the same
// permutation, and up to 16 cells std::sort is a (stable) insertion sort; for a larger column that holds a tie the
// group replays libstdc++'s introsort (ogn_stdsort.h) so that tied cells get the ranks the reference gives them.
template <int G>
__device__ __forceinline__ int ienc_rank(unsigned sm, float sum, int hc, int Hz, bool act) {
    int rank = 0;
    bool tie = false;
    for (int j = 0; j < Hz; ++j) {
        const float sj = __shfl_sync(sm, sum, j, G);
        if (sj > sum || (sj == sum && j < hc)) rank++;
        tie |= act && j != hc && sj == sum;
    }
    if (G == 32 && Hz > 16) {
        if (__any_sync(sm, tie)) {
            float key[32];
            int idx[32];
            for (int j = 0; j < Hz; ++j) { key[j] = __shfl_sync(sm, sum, j, G); idx[j] = j; }
            ogn_stdsort::sort_desc(key, idx, Hz);
            for (int i = 0; i < Hz; ++i) if (idx[i] == hc) rank = i;
        }
    }
    return rank;
}

template <int G>
__global__ void __launch_bounds__(kThreads)
k_ienc_step(const __grid_constant__ ogn_ienc_args A, int Hx, int Hy, int Hz, int npacks, int packsPerRow, int learn, float alpha, float gamma,
            float *resources, int *__restrict__ hiddenCs) {
    constexpr int P = Pack<G>::P, GP = Pack<G>::GP;
    const int gl = threadIdx.x & (GP - 1);
    const int pack = (blockIdx.x * kThreads + threadIdx.x) / GP;
    if (pack >= npacks) return;
    const unsigned sm = lane_mask<G>();
    const int ox = pack / packsPerRow, q = pack % packsPerRow;
    const int p = gl / G, hc = gl % G;
    const int oy = q * P + p;
    const bool colValid = oy < Hy, act = colValid && hc < Hz;
    const int hcl = hc < Hz ? hc : Hz - 1;
    const long long col = (long long)ox * Hy + (colValid ? oy : Hy - 1);

    float sum = 0.0f;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[A.vl[v].geo];
        const PackCtx c = pack_ctx(rf, ox, q, p);
        sum -= ienc_distance2<G>(A.vl[v].wf + c.blk + p * Hz + hcl, A.vl[v].in, c, rf, P * Hz, act);
    }
    float bv = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
    int bi = act ? hc : 0x7fffffff;
    group_argmax<G>(sm, bv, bi);
    if (hc == 0 && colValid) hiddenCs[col] = bv > -999999.0f ? bi : 0;
    if (!learn) return;
    const int rank = ienc_rank<G>(sm, sum, hc, Hz, act);
    if (!act) return;
    float *res = resources + col * Hz + hc;
    const float r = *res;
    const float strength = ogn_expf((float)(-rank * rank) * gamma / fmaxf(0.001f, r)) * r;
    *res = r - alpha * strength;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[A.vl[v].geo];
        const PackCtx c = pack_ctx(rf, ox, q, p);
        ienc_hebb<G>(A.vl[v].wf + c.blk + p * Hz + hcl, A.vl[v].in, c, rf, P * Hz, act, strength);
    }
}

__global__ void __launch_bounds__(kThreads)
k_ienc_reconstruct(const __grid_constant__ ogn_ienc_args A, int vli, int Hy, int Hz, const int *__restrict__ hcs) {
    const ogn_rf &rf = A.geo[A.vl[vli].geo];
    const long long n = (long long)rf.Vx * rf.Vy * rf.Vz;
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int vz = (int)(i % rf.Vz);
    const int iy = (int)((i / rf.Vz) % rf.Vy), ix = (int)(i / ((long long)rf.Vz * rf.Vy));
    const int hx0 = __ldg(rf.rlox + ix), hx1 = __ldg(rf.rhix + ix), hy0 = __ldg(rf.rloy + iy), hy1 = __ldg(rf.rhiy + iy);
    float sum = 0.0f;
    int count = 0;
    for (int ox = hx0; ox <= hx1; ++ox)
        for (int oy = hy0; oy <= hy1; ++oy) {
            const int q = oy / rf.pf, p = oy % rf.pf;
            const int su = (ix - __ldg(rf.lox + ox)) * __ldg(rf.nyf + q) + (iy - __ldg(rf.loyf + q));
            const int hc = __ldg(hcs + ox * Hy + oy);
            sum += A.vl[vli].wf[f_block(rf, ox, q) + ((long long)(su * rf.Vz + vz) * rf.pf + p) * Hz + hc];
            count++;
        }
    A.vl[vli].recon[i] = sum / (float)max(1, count);
}

#include "ogn_vec.cuh"

// ---------------------------------------------------------------------------------------------------------------
// ImageEncoder step with the pack's weight block staged in shared memory. In the F layout the weights of one column pack
// are ONE contiguous block (union slots x Vz rows of P*Hz floats; 31 KB for a radius-4 RGB field), every slot of which the
// dense distance reads: one warp owns one pack, an elected lane pulls the block(s) in with cp.async.bulk (completion on an
// mbarrier) while the warp stages the dense input patch, the distance / arg-max / ranking run out of shared memory, the
// Hebbian pass updates the block in place and one bulk store writes it back. DRAM sees every weight once in and once out
// (8 B per weight and learning step instead of the 12 B of the two-pass kernel above), and with ~7 packs resident per SM
// there are >200 KB of copies in flight per SM without any register cost. Sums run in the reference's order (visible layer,
// slot x-major, y, visible cell), so results are bit-identical to k_ienc_step.
struct IencTileLayout { int wOff[OGN_MAX_IENC_VL], inOff[OGN_MAX_IENC_VL], barOff; }; // float offsets into the warp's smem

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(vec::smem_u32(src)), "r"(bytes) : "memory");
}

// n consecutive rows of the staged block (this lane's float of every row, rows RW floats apart) against the staged patch:
// eight rows of loads in flight before the first add, adds in row order (the reference's order). RW32: the row is one
// 128-byte line (every power-of-two Hz), which turns the row stride into immediate offsets.
template <bool RW32>
__device__ __forceinline__ float ienc_rows_d2(const float *__restrict__ wp, const float *__restrict__ sp, int n, int rowW, float d2) {
    const int RW = RW32 ? 32 : rowW;
    for (; n >= 8; n -= 8, wp += 8 * RW, sp += 8) {
        float wv[8], iv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { wv[k] = wp[k * RW]; iv[k] = sp[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = iv[k] - wv[k]; d2 += d * d; }
    }
    for (; n > 0; --n, wp += RW, ++sp) { const float d = *sp - *wp; d2 += d * d; }
    return d2;
}
template <bool RW32>
__device__ __forceinline__ void ienc_rows_hebb(float *__restrict__ wp, const float *__restrict__ sp, int n, int rowW, float strength) {
    const int RW = RW32 ? 32 : rowW;
    for (; n >= 8; n -= 8, wp += 8 * RW, sp += 8) {
        float wv[8], iv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { wv[k] = wp[k * RW]; iv[k] = sp[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) wp[k * RW] = wv[k] + strength * (iv[k] - wv[k]);
    }
    for (; n > 0; --n, wp += RW, ++sp) { const float old = *wp; *wp = old + strength * (*sp - old); }
}

template <int G, bool RW32>
__global__ void __launch_bounds__(32)
k_ienc_step_t(const __grid_constant__ ogn_ienc_args A, const __grid_constant__ IencTileLayout L, int Hx, int Hy, int Hz, int packsPerRow,
              int learn, float alpha, float gamma, float *resources, int *__restrict__ hiddenCs, long long *trace, int traceBlocks) {
    constexpr int P = Pack<G>::P;
    // debug timeline (ogn_debug_ienc_trace): lane 0 of the first traceBlocks blocks stamps clock64() at every stage
#define IENC_STAMP(i) do { if (trace && threadIdx.x == 0 && (int)blockIdx.x < traceBlocks) trace[(long long)blockIdx.x * 12 + (i)] = clock64(); } while (0)
    if (trace && threadIdx.x == 0 && (int)blockIdx.x < traceBlocks) {
        unsigned smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        trace[(long long)blockIdx.x * 12 + 10] = smid; trace[(long long)blockIdx.x * 12 + 11] = (long long)gt;
    }
    IENC_STAMP(0);
    float *sm = reinterpret_cast<float *>(vec::ogn_dyn_smem);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.barOff); // one per visible layer
    const int gl = threadIdx.x;
    const int pack = blockIdx.x;
    const unsigned gm = lane_mask<G>();
    const int ox = pack / packsPerRow, q = pack % packsPerRow;
    // a pack occupies GP = G * P lanes (32 for 8 or more cells per column, 16 / 8 / 4 below): the warp's other lanes mirror
    // lanes [0, GP) for the collective operations and own no column
    constexpr int GP = Pack<G>::GP;
    const int gle = gl % GP;
    const int p = gle / G, hc = gle % G;
    const int oy = q * P + p;
    const bool colValid = gl < GP && oy < Hy, act = colValid && hc < Hz;
    const long long col = (long long)ox * Hy + (colValid ? oy : Hy - 1);
    const int rowW = P * Hz;
    const int li = p * Hz + (hc < Hz ? hc : Hz - 1);

    if (gl == 0) {
        for (int v = 0; v < A.nvl; ++v) vec::mbar_init_nofence(bars + v);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo];
            const unsigned tb = (unsigned)(__ldg(rf.nx + ox) * __ldg(rf.nyf + q) * rf.Vz * rowW) * 4u;
            vec::mbar_expect_tx(bars + v, tb);
            if (tb) vec::bulk_g2s(sm + L.wOff[v], A.vl[v].wf + f_block(rf, ox, q), tb, bars + v);
        }
    }
    IENC_STAMP(1);
    // this cell's resource, needed after the ranking: start the load now
    float resv = 0.0f;
    if (learn && act) resv = resources[col * Hz + hc];
    // dense input patch of the union field (a row of the patch is contiguous in the input: y-major, cell-minor)
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[A.vl[v].geo];
        const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), lyU = __ldg(rf.loyf + q), nyU = __ldg(rf.nyf + q);
        const int rowF = nyU * rf.Vz, n = nxv * rowF;
        float *si = sm + L.inOff[v];
        const float *base = A.vl[v].in + (long long)rf.Vz * (lyU + (long long)rf.Vy * lx);
        for (int i0 = 0; i0 < n; i0 += 32 * 8) { // eight independent loads per lane before the first store
            float t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = min(i0 + k * 32 + gl, n - 1);
                t[k] = __ldg(base + (long long)(i / rowF) * rf.Vy * rf.Vz + i % rowF);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (i0 + k * 32 + gl < n) si[i0 + k * 32 + gl] = t[k];
        }
    }
    __syncwarp(); // barriers initialised, patch visible to the warp
    IENC_STAMP(2);

    float sum = 0.0f;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[A.vl[v].geo];
        const PackCtx c = pack_ctx(rf, ox, q, p);
        const float *w = sm + L.wOff[v] + li, *si = sm + L.inOff[v];
        vec::mbar_wait(bars + v, 0);
        if (v == 0) IENC_STAMP(3);
        float d2 = 0.0f;
        if (P == 1) // the union field is the column's field: all rows, front to back
            d2 = ienc_rows_d2<RW32>(w, si, c.nslotsU * rf.Vz, rowW, d2);
        else {      // the column's own rows of every visible x are contiguous
            const int nxv = c.nyU > 0 ? c.nslotsU / c.nyU : 0;
            for (int dx = 0; dx < nxv; ++dx) {
                const int r0 = (dx * c.nyU + c.dlo) * rf.Vz;
                d2 = ienc_rows_d2<RW32>(w + (long long)r0 * rowW, si + r0, c.dn * rf.Vz, rowW, d2);
            }
        }
        sum -= act ? d2 : 0.0f;
    }
    IENC_STAMP(4);
    float bv = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
    int bi = act ? hc : 0x7fffffff;
    group_argmax<G>(gm, bv, bi);
    if (hc == 0 && colValid) hiddenCs[col] = bv > -999999.0f ? bi : 0;
    if (!learn) return;
    const int rank = ienc_rank<G>(gm, sum, hc, Hz, act);
    float strength = 0.0f;
    if (act) {
        strength = ogn_expf((float)(-rank * rank) * gamma / fmaxf(0.001f, resv)) * resv;
        resources[col * Hz + hc] = resv - alpha * strength;
    }
    IENC_STAMP(5);
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_rf &rf = A.geo[A.vl[v].geo];
        const PackCtx c = pack_ctx(rf, ox, q, p);
        float *w = sm + L.wOff[v] + li;
        const float *si = sm + L.inOff[v];
        if (act) {
            if (P == 1)
                ienc_rows_hebb<RW32>(w, si, c.nslotsU * rf.Vz, rowW, strength);
            else {
                const int nxv = c.nyU > 0 ? c.nslotsU / c.nyU : 0;
                for (int dx = 0; dx < nxv; ++dx) {
                    const int r0 = (dx * c.nyU + c.dlo) * rf.Vz;
                    ienc_rows_hebb<RW32>(w + (long long)r0 * rowW, si + r0, c.dn * rf.Vz, rowW, strength);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // this lane's generic-proxy writes before the bulk store reads them
        __syncwarp();
        const unsigned tb = (unsigned)(c.nslotsU * rf.Vz * rowW) * 4u;
        if (gl == 0 && tb) bulk_s2g(A.vl[v].wf + c.blk, sm + L.wOff[v], tb);
    }
    IENC_STAMP(6);
    if (gl == 0) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the stores' reads
    }
    IENC_STAMP(7);
#undef IENC_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
// Weight (re)layout, init and checks. One block per hidden column; threads walk the column's CSR-order elements
// e = (hz * nslots + s) * Vz + vz and compute the element's position in the packed device layouts.
struct ColMap {
    int nslots, nyv, ly, Vz, Hz;
    long long csr;        // CSR block offset
    long long fBase;      // F: pack block offset + p*Hz;        element = fBase + ((dx*nyU + dy + fdy)*Vz + vz)*pf*Hz + hz
    int nyU, fdy, fRow;
    long long rBase;      // R: column block offset;             element = rBase + (((dx*ngr + (ly+dy)/pr - ly/pr)*Hz + hz)*pr + (ly+dy)%pr)*Vz + vz
    int ngr, pr;
};
__device__ __forceinline__ ColMap col_map(const ogn_rf &rf, int ox, int oy) {
    ColMap m;
    m.nyv = rf.ny[oy]; m.nslots = rf.nx[ox] * m.nyv; m.ly = rf.loy[oy]; m.Vz = rf.Vz; m.Hz = rf.Hz;
    m.csr = csr_block(rf, ox, oy);
    const int q = oy / rf.pf, p = oy % rf.pf;
    m.fBase = f_block(rf, ox, q) + (long long)p * rf.Hz;
    m.nyU = rf.nyf[q]; m.fdy = m.ly - rf.loyf[q]; m.fRow = rf.pf * rf.Hz;
    m.rBase = r_block(rf, ox, oy);
    m.ngr = rf.ngr[oy]; m.pr = rf.pr;
    return m;
}
__device__ __forceinline__ void split_csr(const ColMap &m, long long e, int &hz, int &dx, int &dy, int &vz) {
    vz = (int)(e % m.Vz);
    const long long t = e / m.Vz;
    const int s = (int)(t % m.nslots);
    hz = (int)(t / m.nslots);
    dx = s / m.nyv; dy = s % m.nyv;
}
__device__ __forceinline__ long long f_index(const ColMap &m, int hz, int dx, int dy, int vz) {
    return m.fBase + ((long long)(dx * m.nyU + dy + m.fdy) * m.Vz + vz) * m.fRow + hz;
}
__device__ __forceinline__ long long r_index(const ColMap &m, int hz, int dx, int dy, int vz) {
    const int iy = m.ly + dy;
    return m.rBase + ((((long long)dx * m.ngr + (iy / m.pr - m.ly / m.pr)) * m.Hz + hz) * m.pr + iy % m.pr) * m.Vz + vz;
}

// which: 0 = F, 1 = R
__global__ void k_layout(const __grid_constant__ ogn_rf rf, float *dev, long long devBase, int which, int col0, float *csr,
                         long long csrBase, int toCsr) {
    const int col = col0 + blockIdx.x;
    const ColMap m = col_map(rf, col / rf.Hy, col % rf.Hy);
    const long long n = (long long)m.nslots * m.Vz * m.Hz;
    float *c = csr + (m.csr - csrBase);
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        int hz, dx, dy, vz;
        split_csr(m, e, hz, dx, dy, vz);
        const long long di = (which == 0 ? f_index(m, hz, dx, dy, vz) : r_index(m, hz, dx, dy, vz)) - devBase;
        if (toCsr) c[e] = dev[di];
        else dev[di] = c[e];
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_init_hash(const __grid_constant__ ogn_rf rf, float *wf, long long wfBase, int fx0, int fx1, float *wr, long long wrBase,
                            int rx0, int rx1, int col0, uint64_t key, float lo, float hi) {
    const int col = col0 + blockIdx.x;
    const int ox = col / rf.Hy;
    const ColMap m = col_map(rf, ox, col % rf.Hy);
    const long long n = (long long)m.nslots * m.Vz * m.Hz;
    const bool inF = wf && ox >= fx0 && ox < fx1, inR = wr && ox >= rx0 && ox < rx1;
    const float span = hi - lo;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        int hz, dx, dy, vz;
        split_csr(m, e, hz, dx, dy, vz);
        const uint32_t u24 = (uint32_t)(mix64(key + (uint64_t)(m.csr + e) * 0xD1B54A32D192ED03ull) >> 40);
        const float val = (float)u24 * (1.0f / 16777216.0f) * span + lo;
        if (inF) wf[f_index(m, hz, dx, dy, vz) - wfBase] = val;
        if (inR) wr[r_index(m, hz, dx, dy, vz) - wrBase] = val;
    }
}

__global__ void k_fr_mismatch(const __grid_constant__ ogn_rf rf, const float *wf, long long wfBase, const float *wr, long long wrBase,
                              int col0, unsigned long long *count) {
    const int col = col0 + blockIdx.x;
    const ColMap m = col_map(rf, col / rf.Hy, col % rf.Hy);
    const long long n = (long long)m.nslots * m.Vz * m.Hz;
    unsigned long long bad = 0;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        int hz, dx, dy, vz;
        split_csr(m, e, hz, dx, dy, vz);
        const float a = wf[f_index(m, hz, dx, dy, vz) - wfBase], b = wr[r_index(m, hz, dx, dy, vz) - wrBase];
        if (__float_as_uint(a) != __float_as_uint(b)) bad++;
    }
    if (bad) atomicAdd(count, bad);
}

__global__ void k_test_expf(const float *in, float *out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ogn_expf(in[i]);
}
__global__ void k_test_tanhf(const float *in, float *out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ogn_tanhf(in[i]);
}


// ---------------------------------------------------------------------------------------------------------------
// Peer-memory all-gather of a sharded CSDR (see ogn_peer_allgather in ogn_kernels_api.h). blockIdx.y = peer.
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256)
k_peer_allgather(const __grid_constant__ ogn_peer_table T, long long off, int count, unsigned epoch, unsigned waitMask,
                 unsigned long long timeoutNs, unsigned *errFlag) {
    __shared__ int lastBlock;
    if (T.world > 1 && count > 0) {
        int p = blockIdx.y;
        if (p >= T.rank) ++p;
        const int *src = T.base[T.rank] + off + (long long)T.rank * count;
        int *dst = T.base[p] + off + (long long)T.rank * count;
        const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
        if ((((unsigned long long)src | (unsigned long long)dst) & 15ull) == 0) {
            const int n4 = count >> 2;
            for (int i = tid; i < n4; i += nth) reinterpret_cast<int4 *>(dst)[i] = __ldg(reinterpret_cast<const int4 *>(src) + i);
            for (int i = (n4 << 2) + tid; i < count; i += nth) dst[i] = __ldg(src + i);
        } else
            for (int i = tid; i < count; i += nth) dst[i] = __ldg(src + i);
    }
    __threadfence_system(); // this thread's peer stores are ordered before the arrival flag below
    __syncthreads();
    unsigned *hdr = reinterpret_cast<unsigned *>(T.base[T.rank]);
    if (threadIdx.x == 0) lastBlock = atomicAdd(&hdr[16], 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (!lastBlock) return;
    if (threadIdx.x == 0) hdr[16] = 0;
    const int q = threadIdx.x;
    if (q < T.world && q != T.rank) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned *>(T.base[q]) + T.rank, epoch); // "rank has delivered exchange #epoch" in q's arena
        const unsigned long long t0 = global_ns();
        unsigned long long t1 = t0;
        // every peer gets the slab and the flag; only the ranks in waitMask are waited for (the ones whose rows the kernels
        // queued behind this exchange read). The other slabs are complete once a later exchange that waits for everybody is.
        while (((waitMask >> q) & 1u) && (int)(ld_acquire_sys(hdr + q) - epoch) < 0) {
            t1 = global_ns();
            if (t1 - t0 > timeoutNs) { *errFlag = 1u; break; }
        }
        // diagnostics: the longest wait of this exchange goes to words 18/19 (ns waited for peers, 64-bit total)
        atomicMax(reinterpret_cast<unsigned long long *>(hdr + 20), t1 - t0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(hdr + 18), *mx = reinterpret_cast<unsigned long long *>(hdr + 20);
        *acc += *mx;
        *mx = 0;
    }
}

inline int group_lanes(int z) {
    int g = 1;
    while (g < z && g < 32) g <<= 1;
    return g;
}
inline int launch_status() { return (int)cudaGetLastError(); }
inline int update_grid_blocks() { // resident blocks of the persistent update kernel: 8 blocks of 256 threads per SM
    static int n = 0;
    if (!n) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        n = sms * 8;
    }
    return n;
}

#define OGN_DISPATCH_G(G_, CALL)                 \
    switch (G_) {                                \
    case 1: { constexpr int G = 1; CALL; } break;   \
    case 2: { constexpr int G = 2; CALL; } break;   \
    case 4: { constexpr int G = 4; CALL; } break;   \
    case 8: { constexpr int G = 8; CALL; } break;   \
    case 16: { constexpr int G = 16; CALL; } break; \
    default: { constexpr int G = 32; CALL; } break; \
    }

} // namespace

extern "C" {

const char *ogn_cuda_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

static inline int pack_factor(int g) { return g >= 8 ? 32 / g : 4; }

// The 128-bit kernels (ogn_vec.cuh) apply when a row holds 8, 16 or 32 cells and the slot descriptors fit their bit fields.
static inline bool vec_row(int z) { return z == 8 || z == 16 || z == 32; }
static inline bool vec_geo(const ogn_rf &g) {
    return (long long)g.max_nx * g.max_nyf * g.Vz < (1ll << 20) && g.max_nyf < 0xFFF && g.max_nx > 0 && g.max_nyf > 0;
}
// OGN_KERNELS selects the step kernels: "scalar" = the generic one-cell-per-lane kernels (always used when a row does not
// hold 8/16/32 cells); "vec1" / "vec2" = 128-bit kernels, one octet per pack, 8 / 16 line loads per lane in registers;
// "tma" = vec1 with the input CSDR windows of a block staged in shared memory by the bulk-copy (TMA) engine; "wide" /
// "wide2" = one warp per pack (four octets share a pack's chunks; 8 / 16 loads per lane). Default (unset): large launches
// (bandwidth-bound) take vec1, small launches (latency-bound, at most kSmallLaunchPacks packs x members) take the
// variant OGN_SMALL_VARIANT names (default below), see profiles/README.md.
constexpr int kVarScalar = 0, kVarVec1 = 1, kVarVec2 = 2, kVarTma = 7, kVarWide = 8, kVarWide2 = 10, kAutoVariant = 9;
constexpr int kVarStage = 11; // Predictor only: weight rows staged in shared memory by cp.async (the other kernels run vec1)
constexpr long long kSmallLaunchPacks = 8192; // packs x members at or below which a launch counts as latency-bound
static int parse_variant(const char *e, int dflt) {
    if (!e) return dflt;
    if (!strcmp(e, "scalar")) return kVarScalar;
    if (!strcmp(e, "vec1")) return kVarVec1;
    if (!strcmp(e, "vec2")) return kVarVec2;
    if (!strcmp(e, "tma")) return kVarTma;
    if (!strcmp(e, "wide")) return kVarWide;
    if (!strcmp(e, "wide2")) return kVarWide2;
    if (!strcmp(e, "stage")) return kVarStage;
    return dflt;
}
static int kernel_variant() {
    static int v = -1;
    if (v < 0) v = parse_variant(getenv("OGN_KERNELS"), kAutoVariant);
    return v;
}
static int small_variant() {
    static int v = -1;
    if (v < 0) v = parse_variant(getenv("OGN_SMALL_VARIANT"), kVarWide);
    return v;
}
static int large_variant() {
    static int v = -1;
    if (v < 0) v = parse_variant(getenv("OGN_LARGE_VARIANT"), kVarVec1);
    return v;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int resolve_variant(long long packs) { // the variant one launch over `packs` packs (x members) uses
    const int v = kernel_variant();
    static long long small = -1;
    if (small < 0) small = env_int("OGN_SMALL_PACKS", (int)kSmallLaunchPacks);
    return v == kAutoVariant ? (packs <= small ? small_variant() : large_variant()) : v;
}
static bool is_wide(int variant) { return variant == kVarWide || variant == kVarWide2; }
static int vec_threads(int variant) { // threads per block of the 128-bit kernels (a multiple of 32, at most 256); OGN_VEC_THREADS overrides
    static int env = -1;
    if (env < 0) env = env_int("OGN_VEC_THREADS", 0);
    int v = env ? env : kThreads;
    if (v < 32 || v > kThreads || v % 32) v = kThreads;
    return v;
}
// launches per (kernel: 0 sc_forward, 1 sc_recon, 2 pred_step; variant column: 0 scalar, 1 vec1, 2 vec2, 3 / 4 unused (the
// cp.async ring variants of round 1), 5 tma, 6 wide, 7 wide2) since the library was loaded: lets a test assert which kernel
// a size actually ran (ogn_debug_variant_launches)
static long long g_variantLaunches[3][8];
static long long *g_iencTrace = nullptr; // ogn_debug_ienc_trace
static unsigned long long *g_smallTrace = nullptr; // ogn_debug_small_trace
static int g_iencTraceBlocks = 0;
static void count_variant(int kernel, int variant) {
    const int col = variant == kVarScalar ? 0 : variant == kVarVec1 ? 1 : variant == kVarVec2 ? 2 : variant == kVarStage ? 3 : variant == kVarTma ? 5 : variant == kVarWide ? 6 : 7;
    g_variantLaunches[kernel][col]++;
}

} // extern "C"
// <<<>>> with dynamic shared memory above the 48 KB default opted in per kernel (again whenever a launch needs more than
// any earlier one of the same kernel did)
template <typename... P, typename... A>
static void launch_k(void (*k)(P...), dim3 grid, int threads, size_t smem, cudaStream_t st, A... a) {
    if (smem > 48u * 1024u) {
        static std::mutex mtx;
        static std::map<const void *, size_t> optedIn;
        std::lock_guard<std::mutex> lock(mtx);
        size_t &have = optedIn[(const void *)k];
        if (smem > have) {
            cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            have = smem;
        }
    }
    k<<<grid, threads, smem, st>>>(a...);
}

extern "C" {

#define OGN_DISPATCH_Z(Z_, CALL)                     \
    switch (Z_) {                                    \
    case 8: { constexpr int Z = 8; CALL; } break;    \
    case 16: { constexpr int Z = 16; CALL; } break;  \
    default: { constexpr int Z = 32; CALL; } break;  \
    }
// CALL sees the compile-time constants Z, NCH (register batch depth) and W (lanes per pack)
#define OGN_DISPATCH_VEC(VARIANT_, Z_, CALL)                                                    \
    switch (VARIANT_) {                                                                         \
    case kVarVec2: { constexpr int NCH = 2, W = 8; OGN_DISPATCH_Z(Z_, CALL); } break;           \
    case kVarWide: { constexpr int NCH = 1, W = 32; OGN_DISPATCH_Z(Z_, CALL); } break;          \
    case kVarWide2: { constexpr int NCH = 2, W = 32; OGN_DISPATCH_Z(Z_, CALL); } break;         \
    default: { constexpr int NCH = 1, W = 8; OGN_DISPATCH_Z(Z_, CALL); } break;                 \
    }

// ints per staged CSDR window for a launch whose blocks hold `ppb` packs of one row, or 0 if the geometry cannot be staged:
// window width <= (y extent the ppb packs' union fields span) rounded out to multiples of 4, height <= max_nx. The bound
// uses only scalars of ogn_rf (max_nyf, pf and the grid ratio), so no table has to be read back from the device.
static int tile_ints(const ogn_rf &g, int ppb) {
    if (g.Vy % 4 != 0 || g.max_nx <= 0 || g.max_nyf <= 0) return 0;
    if (!g.tma_ok) return 0; // host-side geometry check: non-empty, monotone union fields
    // union field of pack q starts at project(q * pf) - r (clamped): consecutive packs advance by at most
    // ceil(pf * Vy / Hy) + 1 visible columns, so ppb packs span at most that times (ppb - 1) plus the widest field
    const long long adv = ((long long)g.pf * g.Vy + g.Hy - 1) / g.Hy + 1;
    long long w = adv * (ppb - 1) + g.max_nyf + 6; // + 6: both ends rounded out to multiples of 4
    w = (w + 3) & ~3ll;
    if (w > g.Vy) w = g.Vy; // Vy % 4 == 0
    const long long ints = w * g.max_nx;
    return ints > 0 && ints < (1 << 14) ? (int)ints : 0;
}

int ogn_vec_row_ok(int z) { return vec_row(z) ? 1 : 0; }
int ogn_vec_geo_ok(const ogn_rf *g) { return vec_geo(*g) ? 1 : 0; }
int ogn_sc_update_version(const ogn_learn_args *args) {
    static int forced = -1;
    if (forced < 0) forced = env_int("OGN_SC_UPDATE", 0);
    const bool v2ok = args->geo.Hz % 8 == 0;
    // default: version 1. Version 2 (full 32-byte sector stores into F) was measured on B200 (profiles/README.md): the L2
    // still fills the sector from DRAM (the DRAM-side write atom is 64 bytes), so it only adds the R group reads.
    if (forced == 2) return v2ok ? 2 : 1;
    if (forced == 3) return args->geo.Hz % 16 == 0 ? 3 : (v2ok ? 2 : 1); // groups of 16 hidden cells: full 64-byte stores
    return 1;
}
int ogn_sc_update_batches(void) {
    static int b = -1;
    if (b < 0) { b = env_int("OGN_UPD2_BATCHES", 4); if (b < 1) b = 1; }
    return b;
}
int ogn_sc_update_chunks(const ogn_learn_args *args, int version) {
    const int Vz = args->geo.Vz;
    int items = (32 / (Vz / 4)) * 4;
    if (version == 2) items = ((2 * Vz >= 32 ? 1 : 32 / (2 * Vz)) * 4) * ogn_sc_update_batches();
    if (version == 3) items = ((4 * Vz >= 32 ? 1 : 32 / (4 * Vz)) * 2) * ogn_sc_update_batches();
    return (args->geo.max_cov + items - 1) / items;
}

int ogn_sc_forward(const ogn_fwd_args *args, int Hx, int Hy, int Hz, int x0, int x1, int batch, int *hidden_cs, int *hidden_cs2,
                   void *stream) {
    if (args->nvl > OGN_MAX_VL) return (int)cudaErrorInvalidValue;
    if (x1 <= x0 || batch <= 0 || Hy <= 0) return 0;
    const int g = group_lanes(Hz), P = pack_factor(g), ppr = (Hy + P - 1) / P, npacks = (x1 - x0) * ppr;
    bool vecOk = vec_row(Hz) && kernel_variant() > 0;
    for (int v = 0; v < args->nvl; v++) {
        if (args->geo[args->vl[v].geo].pf != P) return (int)cudaErrorInvalidValue;
        vecOk = vecOk && vec_geo(args->geo[args->vl[v].geo]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (vecOk) {
        int var = resolve_variant((long long)npacks * batch);
        if (var == kVarStage) var = kVarVec1; // only the Predictor has the staged-rows variant
        const int vt = vec_threads(var);
        if (var == kVarTma) { // staged CSDR windows: every visible layer must be eligible, else plain vec1
            int ti = 0;
            bool ok = ppr % (vt / 8) == 0 && args->nvl <= 8;
            for (int v = 0; v < args->nvl && ok; v++) {
                const int t = tile_ints(args->geo[args->vl[v].geo], vt / 8);
                ok = t > 0;
                ti = t > ti ? t : ti;
            }
            const size_t smem = (size_t)args->nvl * ti * sizeof(int) + 16;
            if (ok && smem <= 96u * 1024u) {
                count_variant(0, kVarTma);
                dim3 grid((unsigned)(((long long)npacks * 8 + vt - 1) / vt), (unsigned)batch);
                OGN_DISPATCH_Z(Hz, (launch_k(vec::k_sc_forward_t<Z, 1>, grid, vt, smem, st, *args, Hx, Hy, x0, npacks, ppr, ti, hidden_cs,
                                             hidden_cs2)));
                return launch_status();
            }
            var = kVarVec1;
        }
        count_variant(0, var);
        const int lanes = is_wide(var) ? 32 : 8;
        dim3 grid((unsigned)(((long long)npacks * lanes + vt - 1) / vt), (unsigned)batch);
        OGN_DISPATCH_VEC(var, Hz, (launch_k(vec::k_sc_forward_v<Z, NCH, W>, grid, vt, 0, st, *args, Hx, Hy, x0, npacks, ppr, hidden_cs,
                                            hidden_cs2)));
        return launch_status();
    }
    count_variant(0, 0);
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads), (unsigned)batch);
    OGN_DISPATCH_G(g, (k_sc_forward<G><<<grid, kThreads, 0, st>>>(*args, Hx, Hy, Hz, x0, npacks, ppr, hidden_cs, hidden_cs2)));
    return launch_status();
}

int ogn_sc_learn_is_split(const ogn_learn_args *args) {
    return vec_row(args->geo.Vz) && kernel_variant() > 0 && args->work && args->work_counters;
}

int ogn_sc_learn(const ogn_learn_args *args, const int *hidden_cs, const int *hidden_cs_prev, int vx0, int vx1, float alpha, int batch,
                 unsigned long long *update_count, int phase, void *stream) {
    if (vx1 <= vx0 || batch <= 0 || args->nvl <= 0) return 0;
    if (args->nvl > OGN_MAX_VL) return (int)cudaErrorInvalidValue;
    const int g = group_lanes(args->geo.Vz), P = pack_factor(g), ppr = (args->geo.Vy + P - 1) / P, npacks = (vx1 - vx0) * ppr;
    if (args->geo.pr != P) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    if (ogn_sc_learn_is_split(args)) {
        const int Vz = args->geo.Vz;
        int var = resolve_variant((long long)npacks * batch * args->nvl);
        if (var == kVarTma || var == kVarStage) var = kVarVec1; // the reconstruction has no staged variant
        const int vt = vec_threads(var);
        const int lanes = is_wide(var) ? 32 : 8;
        dim3 grid((unsigned)(((long long)npacks * lanes + vt - 1) / vt), (unsigned)batch, (unsigned)args->nvl);
        int2 *work = (int2 *)args->work;
        unsigned *cnt = args->work_counters;
        if (phase != OGN_LEARN_UPDATE) {
            count_variant(1, var);
            OGN_DISPATCH_VEC(var, Vz, (launch_k(vec::k_sc_recon_v<Z, NCH, W>, grid, vt, 0, st, *args, hidden_cs, vx0, npacks, ppr, work, cnt)));
            int e = launch_status();
            if (e || phase == OGN_LEARN_RECONSTRUCT) return e;
        }
        // persistent grid over the work units (chunks of cover items of a listed column): enough warps to fill the machine,
        // never more than one per possible unit
        const int ver = ogn_sc_update_version(args);
        const int chunks = ogn_sc_update_chunks(args, ver);
        if (chunks <= 0) return 0;
        long long maxUnits = (long long)(vx1 - vx0) * args->geo.Vy * batch * args->nvl * chunks;
        long long blocks = (maxUnits + (kThreads / 32) - 1) / (kThreads / 32);
        if (blocks > update_grid_blocks()) blocks = update_grid_blocks();
        if (ver == 3) {
            OGN_DISPATCH_Z(Vz, (vec::k_sc_update_v<Z, 3><<<(unsigned)blocks, kThreads, 0, st>>>(*args, hidden_cs, hidden_cs_prev, work, cnt,
                                                                                             cnt + 1, alpha, update_count, chunks,
                                                                                             ogn_sc_update_batches())));
        } else if (ver == 2) {
            OGN_DISPATCH_Z(Vz, (vec::k_sc_update_v<Z, 2><<<(unsigned)blocks, kThreads, 0, st>>>(*args, hidden_cs, hidden_cs_prev, work, cnt,
                                                                                             cnt + 1, alpha, update_count, chunks,
                                                                                             ogn_sc_update_batches())));
        } else {
            OGN_DISPATCH_Z(Vz, (vec::k_sc_update_v<Z, 1><<<(unsigned)blocks, kThreads, 0, st>>>(*args, hidden_cs, hidden_cs_prev, work, cnt,
                                                                                             cnt + 1, alpha, update_count, chunks, 1)));
        }
        return launch_status();
    }
    if (phase == OGN_LEARN_UPDATE) return 0;
    count_variant(1, 0);
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads), (unsigned)batch, (unsigned)args->nvl);
    OGN_DISPATCH_G(g, (k_sc_learn<G><<<grid, kThreads, 0, st>>>(*args, hidden_cs, hidden_cs_prev, vx0, npacks, ppr, alpha, update_count)));
    return launch_status();
}

int ogn_pred_step(const ogn_pred_args *args, int Hx, int Hy, int Hz, int x0, int x1, int batch, int learn, int activate,
                  const int *target_cs, float alpha, float *activations, int *hidden_cs, void *stream) {
    if (args->nvl > OGN_MAX_PRED_VL) return (int)cudaErrorInvalidValue;
    if (x1 <= x0 || batch <= 0 || Hy <= 0) return 0;
    const int g = group_lanes(Hz), P = pack_factor(g), ppr = (Hy + P - 1) / P, npacks = (x1 - x0) * ppr;
    bool vecOk = vec_row(Hz) && kernel_variant() > 0;
    for (int v = 0; v < args->nvl; v++) {
        if (args->geo[args->vl[v].geo].pf != P) return (int)cudaErrorInvalidValue;
        vecOk = vecOk && vec_geo(args->geo[args->vl[v].geo]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (vecOk) {
        int var = resolve_variant((long long)npacks * batch);
        const int vt = vec_threads(var);
        if (var == kVarTma) {
            int ti = 0;
            bool ok = ppr % (vt / 8) == 0 && args->nvl <= 4;
            for (int v = 0; v < args->nvl && ok; v++) {
                const int t = tile_ints(args->geo[args->vl[v].geo], vt / 8);
                ok = t > 0;
                ti = t > ti ? t : ti;
            }
            const size_t smem = (size_t)2 * args->nvl * ti * sizeof(int) + 16;
            if (ok && smem <= 96u * 1024u) {
                count_variant(2, kVarTma);
                dim3 grid((unsigned)(((long long)npacks * 8 + vt - 1) / vt), (unsigned)batch);
                OGN_DISPATCH_Z(Hz, (launch_k(vec::k_pred_step_t<Z, 1, 1>, grid, vt, smem, st, *args, Hx, Hy, x0, npacks, ppr, ti, learn,
                                             activate, target_cs, alpha, activations, hidden_cs)));
                return launch_status();
            }
            var = kVarVec1;
        }
        if (var == kVarStage) {
            static const int ns = env_int("OGN_STAGE_CHUNKS", 3); // 8 * ns rows per lane and round: 1..3
            count_variant(2, kVarStage);
            dim3 grid((unsigned)(((long long)npacks * 8 + vt - 1) / vt), (unsigned)batch);
            const size_t smem = (size_t)vt * 8 * (ns >= 3 ? 3 : ns == 2 ? 2 : 1) * sizeof(float4);
            if (ns >= 3) { OGN_DISPATCH_Z(Hz, (launch_k(vec::k_pred_step_a<Z, 3>, grid, vt, smem, st, *args, Hx, Hy, x0, npacks, ppr, learn, activate, target_cs, alpha, activations, hidden_cs))); }
            else if (ns == 2) { OGN_DISPATCH_Z(Hz, (launch_k(vec::k_pred_step_a<Z, 2>, grid, vt, smem, st, *args, Hx, Hy, x0, npacks, ppr, learn, activate, target_cs, alpha, activations, hidden_cs))); }
            else { OGN_DISPATCH_Z(Hz, (launch_k(vec::k_pred_step_a<Z, 1>, grid, vt, smem, st, *args, Hx, Hy, x0, npacks, ppr, learn, activate, target_cs, alpha, activations, hidden_cs))); }
            return launch_status();
        }
        count_variant(2, var);
        const int lanes = is_wide(var) ? 32 : 8;
        dim3 grid((unsigned)(((long long)npacks * lanes + vt - 1) / vt), (unsigned)batch);
        OGN_DISPATCH_VEC(var, Hz, (launch_k(vec::k_pred_step_v<Z, NCH, 1, W>, grid, vt, 0, st, *args, Hx, Hy, x0, npacks, ppr, learn,
                                            activate, target_cs, alpha, activations, hidden_cs)));
        return launch_status();
    }
    count_variant(2, 0);
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads), (unsigned)batch);
    OGN_DISPATCH_G(g, (k_pred_step<G><<<grid, kThreads, 0, st>>>(*args, Hx, Hy, Hz, x0, npacks, ppr, learn, activate, target_cs, alpha,
                                                              activations, hidden_cs)));
    return launch_status();
}

int ogn_image_encoder_step(const ogn_ienc_args *args, int Hx, int Hy, int Hz, int learn, float alpha, float gamma, float *resources, int *hidden_cs,
                  void *stream) {
    if (args->nvl <= 0 || args->nvl > OGN_MAX_IENC_VL || Hz > 32 || Hz <= 0) return (int)cudaErrorInvalidValue;
    const int g = group_lanes(Hz), P = pack_factor(g), ppr = (Hy + P - 1) / P, npacks = Hx * ppr;
    for (int v = 0; v < args->nvl; v++)
        if (args->geo[args->vl[v].geo].pf != P) return (int)cudaErrorInvalidValue;
    if (npacks <= 0) return 0;
    // tile-staged kernel when the largest pack's blocks and input patches fit the shared memory of an SM (OGN_IENC_TILE=0: never)
    static const int useTile = env_int("OGN_IENC_TILE", 1);
    IencTileLayout L;
    int floats = 0;
    for (int v = 0; v < args->nvl; v++) {
        const ogn_rf &rf = args->geo[args->vl[v].geo];
        L.wOff[v] = floats;
        floats += rf.max_nx * rf.max_nyf * rf.Vz * P * Hz; // multiple of 4 floats when P * Hz is; checked below
    }
    for (int v = 0; v < args->nvl; v++) {
        const ogn_rf &rf = args->geo[args->vl[v].geo];
        L.inOff[v] = floats;
        floats += (rf.max_nx * rf.max_nyf * rf.Vz + 3) & ~3;
    }
    L.barOff = floats;
    const size_t smem = (size_t)floats * 4 + 8 * (size_t)OGN_MAX_IENC_VL;
    if (useTile && (P * Hz) % 4 == 0 && smem <= 200 * 1024) {
        const cudaStream_t st = (cudaStream_t)stream;
        const dim3 grid((unsigned)npacks);
        if (P * Hz == 32) {
            OGN_DISPATCH_G(g, (launch_k(k_ienc_step_t<G, true>, grid, 32, smem, st, *args, L, Hx, Hy, Hz, ppr, learn, alpha, gamma, resources, hidden_cs,
                                        g_iencTrace, g_iencTraceBlocks)));
        } else {
            OGN_DISPATCH_G(g, (launch_k(k_ienc_step_t<G, false>, grid, 32, smem, st, *args, L, Hx, Hy, Hz, ppr, learn, alpha, gamma, resources, hidden_cs,
                                        g_iencTrace, g_iencTraceBlocks)));
        }
        return launch_status();
    }
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads));
    OGN_DISPATCH_G(g, (k_ienc_step<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(*args, Hx, Hy, Hz, npacks, ppr, learn, alpha, gamma, resources,
                                                                                hidden_cs)));
    return launch_status();
}

int ogn_image_encoder_reconstruct(const ogn_ienc_args *args, int vli, int Hx, int Hy, int Hz, const int *hidden_cs, void *stream) {
    if (vli < 0 || vli >= args->nvl) return (int)cudaErrorInvalidValue;
    const ogn_rf &rf = args->geo[args->vl[vli].geo];
    const long long n = (long long)rf.Vx * rf.Vy * rf.Vz;
    if (n <= 0) return 0;
    k_ienc_reconstruct<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(*args, vli, Hy, Hz, hidden_cs);
    return launch_status();
}

int ogn_hier_small_max_blocks(int threads) {
    int dev = 0, sms = 0, per = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, vec::k_hier_step_small<32, vec::CsCg>, threads, 48 * 1024) != cudaSuccess) return 0;
    return sms * per;
}

int ogn_hier_step_small(const ogn_small_phase *phases, int nphases, const ogn_small_inputs *inputs, int members, int blocks_per_member,
                        int threads, int wide, unsigned *counters, void *stream) {
    if (nphases <= 0 || members <= 0) return 0;
    if (blocks_per_member < 1 || threads < 32 || threads > 512 || threads % 32) return (int)cudaErrorInvalidValue;
    dim3 grid((unsigned)blocks_per_member, (unsigned)members);
    int cpm = OGN_SMALL_COUNTERS;
    ogn_small_inputs in;
    memset(&in, 0, sizeof(in));
    if (inputs) in = *inputs;
    // the phase array goes to shared memory when it fits the default 48 KB (a 4-layer step is ~35 KB)
    int phaseBytes = nphases * (int)sizeof(ogn_small_phase);
    if (phaseBytes > 48 * 1024) phaseBytes = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (blocks_per_member == 1) { // one block owns a member: CSDRs written by the block are read back through its own L1
        if (wide) vec::k_hier_step_small<32, vec::CsSm><<<grid, threads, phaseBytes, st>>>(phases, nphases, in, counters, cpm, phaseBytes, g_smallTrace);
        else vec::k_hier_step_small<8, vec::CsSm><<<grid, threads, phaseBytes, st>>>(phases, nphases, in, counters, cpm, phaseBytes, g_smallTrace);
        return launch_status();
    }
    unsigned long long *trace = g_smallTrace;
    void *kargs[] = {(void *)&phases, (void *)&nphases, (void *)&in, (void *)&counters, (void *)&cpm, (void *)&phaseBytes, (void *)&trace};
    const void *fn = wide ? (const void *)vec::k_hier_step_small<32, vec::CsCg> : (const void *)vec::k_hier_step_small<8, vec::CsCg>;
    return (int)cudaLaunchCooperativeKernel(fn, grid, dim3((unsigned)threads), kargs, (size_t)phaseBytes, st);
}

int ogn_actor_forward(const ogn_actor_args *args, int Hx, int Hy, int Hz, const float *u01, float *hidden_values,
                      float *hidden_activations, int *hidden_cs, void *stream) {
    if (args->nvl > OGN_MAX_PRED_VL) return (int)cudaErrorInvalidValue;
    static const int rowsKernel = env_int("OGN_ACTOR_ROWS", 1);
    if (rowsKernel && Hz <= 31) { // one block per action column, one warp per weight row
        k_actor_forward_rows<<<(unsigned)(Hx * Hy), 32 * (1 + Hz), 0, (cudaStream_t)stream>>>(*args, Hx, Hy, Hz, u01, hidden_values,
                                                                                         hidden_activations, hidden_cs);
        return launch_status();
    }
    const int g = group_lanes(Hz), P = pack_factor(g), ppr = (Hy + P - 1) / P, npacks = Hx * ppr;
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads));
    OGN_DISPATCH_G(g, (k_actor_forward<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(*args, Hx, Hy, Hz, npacks, ppr, u01, hidden_values,
                                                                                    hidden_activations, hidden_cs)));
    return launch_status();
}

int ogn_actor_learn(const ogn_actor_args *args, const ogn_actor_batch *batch, int Hx, int Hy, int Hz, const float *hidden_values,
                    float alpha, float beta, int mimic, float *hidden_activations, void *stream) {
    if (args->nvl > OGN_MAX_PRED_VL || batch->n < 0 || batch->n > OGN_MAX_ACTOR_SAMPLES) return (int)cudaErrorInvalidValue;
    if (batch->n == 0) return 0;
    static const int rowsKernel = env_int("OGN_ACTOR_ROWS", 1);
    if (rowsKernel && Hz <= 31) { // one block per action column, one warp per weight row
        k_actor_learn_rows<<<(unsigned)(Hx * Hy), 32 * (1 + Hz), 0, (cudaStream_t)stream>>>(*args, *batch, Hx, Hy, Hz, hidden_values, alpha, beta,
                                                                                       mimic, hidden_activations);
        return launch_status();
    }
    const int g = group_lanes(Hz), P = pack_factor(g), ppr = (Hy + P - 1) / P, npacks = Hx * ppr;
    dim3 grid((unsigned)(((long long)npacks * g * P + kThreads - 1) / kThreads));
    OGN_DISPATCH_G(g, (k_actor_learn<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(*args, *batch, Hx, Hy, Hz, npacks, ppr, hidden_values,
                                                                                  alpha, beta, mimic, hidden_activations)));
    return launch_status();
}

int ogn_peer_allgather(const ogn_peer_table *table, int64_t offset_ints, int count_per_rank, unsigned epoch, unsigned wait_mask,
                       unsigned long long timeout_ns, unsigned *err_flag, void *stream) {
    if (table->world <= 1) return 0;
    if (table->world > OGN_MAX_PEERS || table->rank < 0 || table->rank >= table->world) return (int)cudaErrorInvalidValue;
    int bx = (count_per_rank / 4 + 1023) / 1024; // ~4 x 16-byte stores per thread
    if (bx < 1) bx = 1;
    if (bx > 16) bx = 16;
    dim3 grid((unsigned)bx, (unsigned)(table->world - 1));
    k_peer_allgather<<<grid, 256, 0, (cudaStream_t)stream>>>(*table, offset_ints, count_per_rank, epoch, wait_mask, timeout_ns, err_flag);
    return launch_status();
}

int ogn_weights_from_csr(const ogn_wset *ws, const ogn_rf *rf, int member, int which, int col0, int col1, const float *csr,
                         int64_t csr_base, void *stream) {
    if (col1 <= col0) return 0;
    float *dev = which == 0 ? ws->wf : ws->wr;
    if (!dev) return (int)cudaErrorInvalidValue;
    dev += (long long)member * (which == 0 ? ws->wf_mstride : ws->wr_mstride);
    k_layout<<<col1 - col0, 256, 0, (cudaStream_t)stream>>>(*rf, dev, which == 0 ? ws->wf_base : ws->wr_base, which, col0,
                                                            const_cast<float *>(csr), csr_base, 0);
    return launch_status();
}

int ogn_weights_to_csr(const ogn_wset *ws, const ogn_rf *rf, int member, int which, int col0, int col1, float *csr,
                       int64_t csr_base, void *stream) {
    if (col1 <= col0) return 0;
    float *dev = which == 0 ? ws->wf : ws->wr;
    if (!dev) return (int)cudaErrorInvalidValue;
    dev += (long long)member * (which == 0 ? ws->wf_mstride : ws->wr_mstride);
    k_layout<<<col1 - col0, 256, 0, (cudaStream_t)stream>>>(*rf, dev, which == 0 ? ws->wf_base : ws->wr_base, which, col0, csr,
                                                            csr_base, 1);
    return launch_status();
}

int ogn_weights_init_hash(const ogn_wset *ws, const ogn_rf *rf, int member, uint64_t seed, uint32_t stream_id, float lo, float hi,
                          void *stream) {
    int xa = ws->fx0, xb = ws->fx1;
    if (ws->wr) { xa = ws->rx0 < xa ? ws->rx0 : xa; xb = ws->rx1 > xb ? ws->rx1 : xb; }
    if (xb <= xa) return 0;
    const uint64_t key = seed * 0x9E3779B97F4A7C15ull + (uint64_t)stream_id * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull;
    float *wf = ws->wf ? ws->wf + (long long)member * ws->wf_mstride : nullptr;
    float *wr = ws->wr ? ws->wr + (long long)member * ws->wr_mstride : nullptr;
    k_init_hash<<<(xb - xa) * rf->Hy, 256, 0, (cudaStream_t)stream>>>(*rf, wf, ws->wf_base, ws->fx0, ws->fx1, wr, ws->wr_base, ws->rx0,
                                                                      ws->rx1, xa * rf->Hy, key, lo, hi);
    return launch_status();
}

int ogn_weights_fr_mismatch(const ogn_wset *ws, const ogn_rf *rf, int member, unsigned long long *count_dev, void *stream) {
    if (!ws->wf || !ws->wr) return (int)cudaErrorInvalidValue;
    const int xa = ws->fx0 > ws->rx0 ? ws->fx0 : ws->rx0, xb = ws->fx1 < ws->rx1 ? ws->fx1 : ws->rx1;
    if (xb <= xa) return 0;
    k_fr_mismatch<<<(xb - xa) * rf->Hy, 256, 0, (cudaStream_t)stream>>>(*rf, ws->wf + (long long)member * ws->wf_mstride, ws->wf_base,
                                                                        ws->wr + (long long)member * ws->wr_mstride, ws->wr_base,
                                                                        xa * rf->Hy, count_dev);
    return launch_status();
}

// Timeline of the single-launch step: (1, NULL) arms a device buffer of 2 x 256 x 3 globaltimer stamps (blocks 0 and nb - 1 of
// member 0; per phase: start, end of the phase's work, end of its barrier), overwritten by every launch; (0, out) copies the
// 1536 words to the host and disarms. Returns 0 or a CUDA error.
int ogn_debug_small_trace(int arm, unsigned long long *out) {
    if (arm) {
        if (!g_smallTrace) {
            cudaError_t e = cudaMalloc(&g_smallTrace, 1536 * sizeof(unsigned long long));
            if (e != cudaSuccess) return (int)e;
        }
        return (int)cudaMemset(g_smallTrace, 0, 1536 * sizeof(unsigned long long));
    }
    if (!g_smallTrace) return (int)cudaErrorInvalidValue;
    cudaDeviceSynchronize();
    cudaError_t e = out ? cudaMemcpy(out, g_smallTrace, 1536 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) : cudaSuccess;
    cudaFree(g_smallTrace);
    g_smallTrace = nullptr;
    return (int)e;
}

// Timeline of the tile-staged ImageEncoder kernel: (blocks, NULL) arms a device buffer of 12 words per block (stamps 0..7 =
// clock64 at start / copies queued / patch staged / block landed / distances done / ranked / updated / stored, 10 = SM id,
// 11 = globaltimer at start); (blocks, out) copies it to the host and disarms. Returns 0 or a CUDA error.
int ogn_debug_ienc_trace(int blocks, long long *out) {
    if (!out) {
        if (g_iencTrace) cudaFree(g_iencTrace);
        g_iencTrace = nullptr; g_iencTraceBlocks = 0;
        if (blocks <= 0) return 0;
        cudaError_t e = cudaMalloc(&g_iencTrace, (size_t)blocks * 12 * sizeof(long long));
        if (e != cudaSuccess) return (int)e;
        cudaMemset(g_iencTrace, 0, (size_t)blocks * 12 * sizeof(long long));
        g_iencTraceBlocks = blocks;
        return 0;
    }
    if (!g_iencTrace) return (int)cudaErrorInvalidValue;
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpy(out, g_iencTrace, (size_t)(blocks < g_iencTraceBlocks ? blocks : g_iencTraceBlocks) * 12 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(g_iencTrace);
    g_iencTrace = nullptr; g_iencTraceBlocks = 0;
    return (int)e;
}

void ogn_debug_variant_launches(long long out[24]) { memcpy(out, g_variantLaunches, sizeof(g_variantLaunches)); }

int ogn_test_expf(const float *in, float *out, int64_t n, void *stream) {
    if (n <= 0) return 0;
    k_test_expf<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    return launch_status();
}
int ogn_test_tanhf(const float *in, float *out, int64_t n, void *stream) {
    if (n <= 0) return 0;
    k_test_tanhf<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    return launch_status();
}

} // extern "C"
