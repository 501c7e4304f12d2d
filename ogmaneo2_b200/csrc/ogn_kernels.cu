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

#include "../../include/ogn_kernels_api.h"
#include "ogn_libm.h"

namespace {

constexpr int kThreads = 256;

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}

// Offset of hidden column (ox,oy)'s weight block (identical in the CSR, F and R orders).
__device__ __forceinline__ long long col_block(const ogn_rf &rf, int ox, int oy) {
    return (long long)rf.Vz * rf.Hz *
           ((long long)__ldg(rf.px + ox) * rf.sy + (long long)__ldg(rf.nx + ox) * (long long)__ldg(rf.py + oy));
}

// (value, index) max over a lane group; the lowest index wins ties. Values must be NaN-free.
template <int G>
__device__ __forceinline__ void group_argmax(unsigned m, float &v, int &i) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(m, v, o, G);
        int oi = __shfl_xor_sync(m, i, o, G);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// Σ over the receptive field of W_F[col][slot][cs[visible column of slot]][hc] for this lane's hc, in slot order
// (SparseMatrix::multiplyOHVs, SparseMatrix.cpp:260-276). kReadOnly selects the non-coherent load path.
template <int G, bool kReadOnly>
__device__ __forceinline__ float gather_row(unsigned gm, int gl, bool act, const float *w, const int *cs, int lx, int ly,
                                            int nyv, int nslots, int Vy, int Vz, int Hz) {
    float p = 0.0f;
    for (int ch = 0; ch < nslots; ch += G) {
        const int sm = ch + gl;
        int cm = 0;
        if (sm < nslots) cm = __ldg(cs + (lx + sm / nyv) * Vy + ly + sm % nyv);
        float wv[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int cj = __shfl_sync(gm, cm, j, G);
            wv[j] = 0.0f;
            if (act && ch + j < nslots) {
                const float *a = w + ((long long)(ch + j) * Vz + cj) * Hz;
                wv[j] = kReadOnly ? __ldg(a) : *a;
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j)
            if (ch + j < nslots) p += wv[j];
    }
    return p;
}

// W_F[col][slot][cs[...]][hc] += delta over the receptive field (SparseMatrix::deltaOHVs, SparseMatrix.cpp:488-501).
template <int G>
__device__ __forceinline__ void delta_row(unsigned gm, int gl, bool act, float *w, const int *cs, float delta, int lx, int ly,
                                          int nyv, int nslots, int Vy, int Vz, int Hz) {
    for (int ch = 0; ch < nslots; ch += G) {
        const int sm = ch + gl;
        int cm = 0;
        if (sm < nslots) cm = __ldg(cs + (lx + sm / nyv) * Vy + ly + sm % nyv);
        float wv[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int cj = __shfl_sync(gm, cm, j, G);
            wv[j] = 0.0f;
            if (act && ch + j < nslots) wv[j] = w[((long long)(ch + j) * Vz + cj) * Hz];
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int cj = __shfl_sync(gm, cm, j, G);
            if (act && ch + j < nslots) w[((long long)(ch + j) * Vz + cj) * Hz] = wv[j] + delta;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2: SparseCoder::forward (SparseCoder.cpp:13-43)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_sc_forward(const ogn_wset *__restrict__ wsets, int nvl, ogn_cs_args cs, int Hx, int Hy, int Hz, int x0, int ncols,
             int *__restrict__ out0, int *__restrict__ out1) {
    const int gl = threadIdx.x & (G - 1);
    const int group = (blockIdx.x * kThreads + threadIdx.x) / G;
    if (group >= ncols) return; // whole groups leave together
    const unsigned gm = group_mask<G>();
    const int member = blockIdx.y;
    const int ox = x0 + group / Hy, oy = group % Hy;

    int bestIdx = 0;
    float bestVal = -999999.0f;
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + gl;
        const bool act = hc < Hz;
        float sum = 0.0f;
        for (int v = 0; v < nvl; ++v) {
            const ogn_wset ws = wsets[v];
            const ogn_rf rf = *ws.rf;
            const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
            const float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base + (long long)member * ws.wf_mstride) + hc;
            const int *c = cs.cs[v] + (long long)member * rf.Vx * rf.Vy;
            sum += gather_row<G, true>(gm, gl, act, w, c, lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, rf.Hz);
        }
        // sequential rule: `if (sum > max)` from (idx 0, -999999); NaN never wins
        float v = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
        int i = act ? hc : 0x7fffffff;
        group_argmax<G>(gm, v, i);
        if (v > bestVal) { bestVal = v; bestIdx = i; }
    }
    if (gl == 0) {
        const long long o = (long long)member * Hx * Hy + ox * Hy + oy;
        out0[o] = bestIdx;
        if (out1) out1[o] = bestIdx;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: SparseCoder::learn (SparseCoder.cpp:45-83). One lane group per visible column, lane = visible cell vc.
struct CoverItem { long long offR; long long offF; int flags; }; // flags: bit0 changed, bit1 F copy stored here

template <int G>
__device__ __forceinline__ CoverItem cover_item(const ogn_wset &ws, const ogn_rf &rf, int k, int ncov, int ncy, int hx0, int hy0,
                                                int vx, int vy, const int *hcs, const int *hcsPrev) {
    CoverItem it;
    it.offR = 0; it.offF = 0; it.flags = 0;
    if (k < ncov) {
        const int ox = hx0 + k / ncy, oy = hy0 + k % ncy;
        const int h = ox * rf.Hy + oy;
        const int hc = __ldg(hcs + h);
        const int s = (vx - __ldg(rf.lox + ox)) * __ldg(rf.ny + oy) + (vy - __ldg(rf.loy + oy));
        const long long blk = col_block(rf, ox, oy);
        it.offR = blk - ws.wr_base + ((long long)s * rf.Hz + hc) * rf.Vz;
        it.offF = blk - ws.wf_base + (long long)s * rf.Vz * rf.Hz + hc;
        it.flags = (hc != __ldg(hcsPrev + h) ? 1 : 0) | ((ox >= ws.fx0 && ox < ws.fx1) ? 2 : 0);
    }
    return it;
}

template <int G>
__global__ void __launch_bounds__(kThreads)
k_sc_learn(const ogn_wset *__restrict__ wsets, int v0, ogn_cs_args cs, const int *__restrict__ hcsAll,
           const int *__restrict__ hcsPrevAll, int vx0, int ncols, float alpha, unsigned long long *updCount) {
    const int gl = threadIdx.x & (G - 1);
    const int group = (blockIdx.x * kThreads + threadIdx.x) / G;
    if (group >= ncols) return;
    const unsigned gm = group_mask<G>();
    const int member = blockIdx.y;
    const int v = v0 + blockIdx.z;
    const ogn_wset ws = wsets[v];
    const ogn_rf rf = *ws.rf;
    const int Vz = rf.Vz;
    const int vx = vx0 + group / rf.Vy, vy = group % rf.Vy;
    const long long vcol = (long long)member * rf.Vx * rf.Vy + vx * rf.Vy + vy;
    const int target = __ldg(cs.cs[v] + vcol);
    const int *hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
    const int *hcsPrev = hcsPrevAll + (long long)member * rf.Hx * rf.Hy;
    float *wr = ws.wr + (long long)member * ws.wr_mstride;
    float *wf = ws.wf + (long long)member * ws.wf_mstride;
    float *recon = ws.recon + vcol * Vz;

    const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx), hy0 = __ldg(rf.rloy + vy), hy1 = __ldg(rf.rhiy + vy);
    const int ncx = max(0, hx1 - hx0 + 1), ncy = max(0, hy1 - hy0 + 1);
    const int ncov = ncx * ncy; // == countT(visibleIndex) / hiddenSize.z

    // reconstruction + arg-max with the reference's `sum > max || maxIndex == -1` rule
    int bestIdx = -1;
    float bestVal = -999999.0f;
    float myRecon = 0.0f; // recon of vc = gl when Vz <= G
    for (int vc0 = 0; vc0 < Vz; vc0 += G) {
        const int vc = vc0 + gl;
        const bool act = vc < Vz;
        float sum = 0.0f;
        for (int ch = 0; ch < ncov; ch += G) {
            const CoverItem mine = cover_item<G>(ws, rf, ch + gl, ncov, ncy, hx0, hy0, vx, vy, hcs, hcsPrev);
            float wv[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const long long off = __shfl_sync(gm, mine.offR, j, G);
                wv[j] = 0.0f;
                if (act && ch + j < ncov) wv[j] = wr[off + vc];
            }
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (ch + j < ncov) sum += wv[j];
        }
        sum = sum / (float)ncov;
        if (act) recon[vc] = sum;
        if (vc0 == 0) myRecon = sum;
        // chunk arg-max: element 0 of the whole column is always taken first, NaN elsewhere never wins
        float val;
        if (!act) val = -INFINITY;
        else if (vc == 0) val = (sum != sum) ? INFINITY : sum;
        else val = (sum != sum) ? -INFINITY : sum;
        int idx = act ? vc : 0x7fffffff;
        group_argmax<G>(gm, val, idx);
        if (bestIdx == -1 || val > bestVal) { bestVal = val; bestIdx = idx; }
    }
    if (bestIdx == target) return;

    // delta rule on the weights of hidden columns whose state changed (deltaChangedOHVsT, SparseMatrix.cpp:572-590)
    int nupd = 0;
    for (int vc0 = 0; vc0 < Vz; vc0 += G) {
        const int vc = vc0 + gl;
        const bool act = vc < Vz;
        float r = myRecon;
        if (vc0 != 0 && act) r = recon[vc];
        const float delta = alpha * ((vc == target ? 1.0f : 0.0f) - ogn_expf(r));
        for (int ch = 0; ch < ncov; ch += G) {
            const CoverItem mine = cover_item<G>(ws, rf, ch + gl, ncov, ncy, hx0, hy0, vx, vy, hcs, hcsPrev);
#pragma unroll 4
            for (int j = 0; j < G; ++j) {
                const long long offR = __shfl_sync(gm, mine.offR, j, G);
                const long long offF = __shfl_sync(gm, mine.offF, j, G);
                const int fl = __shfl_sync(gm, mine.flags, j, G);
                if (ch + j < ncov && (fl & 1)) {
                    if (act) {
                        const float nw = wr[offR + vc] + delta;
                        wr[offR + vc] = nw;
                        if (fl & 2) wf[offF + (long long)vc * rf.Hz] = nw;
                    }
                    if (vc0 == 0) nupd++;
                }
            }
        }
    }
    if (updCount && gl == 0 && nupd) atomicAdd(updCount, (unsigned long long)nupd);
}

// ---------------------------------------------------------------------------------------------------------------
// K6+K7+K8: Predictor::learn (Predictor.cpp:49-70), Predictor::forward (:13-47), inputCs -> inputCsPrev (:118-126)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_pred_step(const ogn_wset *__restrict__ wsets, int nvl, ogn_pred_args a, int Hx, int Hy, int Hz, int x0, int ncols, int learn,
            int activate, const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    const int member = blockIdx.y;
    // K8: every thread of the grid helps copying the inputs into the *other* inputCsPrev buffer (never read here)
    if (activate) {
        const int tid = blockIdx.x * kThreads + threadIdx.x, nth = gridDim.x * kThreads;
        for (int v = 0; v < nvl; ++v) {
            if (!a.cs_prev_out[v]) continue;
            const ogn_rf *rfp = wsets[v].rf;
            const int n = rfp->Vx * rfp->Vy;
            const long long mo = (long long)member * n;
            for (int i = tid; i < n; i += nth) a.cs_prev_out[v][mo + i] = __ldg(a.cs[v] + mo + i);
        }
    }
    const int gl = threadIdx.x & (G - 1);
    const int group = (blockIdx.x * kThreads + threadIdx.x) / G;
    if (group >= ncols) return;
    const unsigned gm = group_mask<G>();
    const int ox = x0 + group / Hy, oy = group % Hy;
    const long long col = (long long)member * Hx * Hy + ox * Hy + oy;

    if (learn) {
        const int t = __ldg(target + col);
        for (int hc0 = 0; hc0 < Hz; hc0 += G) {
            const int hc = hc0 + gl;
            const bool act = hc < Hz;
            float delta = 0.0f;
            if (act) delta = alpha * ((hc == t ? 1.0f : -1.0f) - ogn_tanhf(acts[col * Hz + hc]));
            for (int v = 0; v < nvl; ++v) {
                const ogn_wset ws = wsets[v];
                const ogn_rf rf = *ws.rf;
                const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
                float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base + (long long)member * ws.wf_mstride) + hc;
                const int *c = a.cs_prev[v] + (long long)member * rf.Vx * rf.Vy;
                delta_row<G>(gm, gl, act, w, c, delta, lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, rf.Hz);
            }
        }
    }
    if (!activate) return;

    int bestIdx = 0;
    float bestVal = -999999.0f;
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + gl;
        const bool act = hc < Hz;
        float sum = 0.0f;
        int count = 0;
        for (int v = 0; v < nvl; ++v) {
            const ogn_wset ws = wsets[v];
            const ogn_rf rf = *ws.rf;
            const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
            const float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base + (long long)member * ws.wf_mstride) + hc;
            const int *c = a.cs[v] + (long long)member * rf.Vx * rf.Vy;
            sum += gather_row<G, false>(gm, gl, act, w, c, lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, rf.Hz);
            count += nxv * nyv;
        }
        sum = sum / (float)count;
        if (act) acts[col * Hz + hc] = sum;
        float v = (act && sum > -999999.0f) ? sum : (act ? -999999.0f : -INFINITY);
        int i = act ? hc : 0x7fffffff;
        group_argmax<G>(gm, v, i);
        if (v > bestVal) { bestVal = v; bestIdx = i; }
    }
    if (gl == 0) hiddenCs[col] = bestIdx;
}

// ---------------------------------------------------------------------------------------------------------------
// Actor helpers: value row = the single row of the (Hx,Hy,1) value matrix; all lanes of the group compute it redundantly
// (same addresses -> broadcast loads).
template <int G>
__device__ __forceinline__ float actor_value_sum(unsigned gm, int gl, const ogn_wset *vw, int nvl, const int *const *cs, int ox,
                                                 int oy, int *countOut) {
    float value = 0.0f;
    int count = 0;
    for (int v = 0; v < nvl; ++v) {
        const ogn_wset ws = vw[v];
        const ogn_rf rf = *ws.rf;
        const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
        const float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base);
        value += gather_row<G, false>(gm, gl, true, w, cs[v], lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, 1);
        count += nxv * nyv;
    }
    *countOut = count;
    return value;
}

// action activations of this lane's hc: (Σ_v gather) / count, stored raw; returns them through acts[]
template <int G>
__device__ __forceinline__ void actor_action_acts(unsigned gm, int gl, const ogn_wset *aw, int nvl, const int *const *cs, int ox,
                                                  int oy, int Hz, int count, float *acts /* column base */) {
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + gl;
        const bool act = hc < Hz;
        float sum = 0.0f;
        for (int v = 0; v < nvl; ++v) {
            const ogn_wset ws = aw[v];
            const ogn_rf rf = *ws.rf;
            const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
            const float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base) + hc;
            sum += gather_row<G, false>(gm, gl, act, w, cs[v], lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, rf.Hz);
        }
        sum = sum / (float)count;
        if (act) acts[hc] = sum;
    }
    __syncwarp(gm);
}

// softmax numerators in place (Actor.cpp:57-69 / :159-171): acts[hc] = expf(acts[hc] - max); returns Σ in ascending hc order.
template <int G>
__device__ __forceinline__ float actor_softmax(unsigned gm, int gl, int Hz, float *acts) {
    float m = -999999.0f;
    for (int hc = 0; hc < Hz; ++hc) m = fmaxf(m, __ldcg(acts + hc)); // every lane scans the same Hz values (L2 reads)
    __syncwarp(gm);
    for (int hc = gl; hc < Hz; hc += G) acts[hc] = ogn_expf(__ldcg(acts + hc) - m);
    __syncwarp(gm);
    float total = 0.0f;
    for (int hc = 0; hc < Hz; ++hc) total += __ldcg(acts + hc);
    return total;
}

// K9: Actor::forward (Actor.cpp:13-90)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_actor_forward(const ogn_wset *__restrict__ vw, const ogn_wset *__restrict__ aw, int nvl, ogn_pred_args a, int Hx, int Hy, int Hz,
                const float *__restrict__ u01, float *hiddenValues, float *hiddenActs, int *hiddenCs) {
    const int gl = threadIdx.x & (G - 1);
    const int group = (blockIdx.x * kThreads + threadIdx.x) / G;
    if (group >= Hx * Hy) return;
    const unsigned gm = group_mask<G>();
    const int ox = group / Hy, oy = group % Hy;
    int count;
    const float value = actor_value_sum<G>(gm, gl, vw, nvl, a.cs, ox, oy, &count);
    if (gl == 0) hiddenValues[group] = value / (float)count;
    float *acts = hiddenActs + (long long)group * Hz;
    actor_action_acts<G>(gm, gl, aw, nvl, a.cs, ox, oy, Hz, count, acts);
    const float total = actor_softmax<G>(gm, gl, Hz, acts);
    if (gl == 0) {
        // std::uniform_real_distribution<float>(0, total): canonical * (b - a) + a   (libstdc++ bits/random.h)
        const float cusp = u01[group] * (total - 0.0f) + 0.0f;
        int select = 0;
        float sumSoFar = 0.0f;
        for (int hc = 0; hc < Hz; ++hc) {
            sumSoFar += __ldcg(acts + hc);
            if (sumSoFar >= cusp) { select = hc; break; }
        }
        hiddenCs[group] = select;
    }
}

// K11: Actor::learn (Actor.cpp:92-185)
template <int G>
__global__ void __launch_bounds__(kThreads)
k_actor_learn(const ogn_wset *__restrict__ vw, const ogn_wset *__restrict__ aw, int nvl, ogn_pred_args a, int Hx, int Hy, int Hz,
              const int *__restrict__ targetPrev, const float *__restrict__ valuesPrev, const float *__restrict__ hiddenValues,
              float q, float g, float alpha, float beta, int mimic, float *hiddenActs) {
    const int gl = threadIdx.x & (G - 1);
    const int group = (blockIdx.x * kThreads + threadIdx.x) / G;
    if (group >= Hx * Hy) return;
    const unsigned gm = group_mask<G>();
    const int ox = group / Hy, oy = group % Hy;

    const float newValue = q + g * hiddenValues[group];
    int count;
    float value = actor_value_sum<G>(gm, gl, vw, nvl, a.cs, ox, oy, &count);
    value = value / (float)count;
    const float tdErrorValue = newValue - value;
    const float deltaValue = alpha * tdErrorValue;
    for (int v = 0; v < nvl; ++v) { // value row += deltaValue: slots spread over the lanes (distinct addresses)
        const ogn_wset ws = vw[v];
        const ogn_rf rf = *ws.rf;
        const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
        float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base);
        for (int s = gl; s < nxv * nyv; s += G) {
            const int c = __ldg(a.cs[v] + (lx + s / nyv) * rf.Vy + ly + s % nyv);
            w[(long long)s * rf.Vz + c] += deltaValue;
        }
    }
    const float tdErrorAction = newValue - valuesPrev[group];
    const int target = targetPrev[group];
    float *acts = hiddenActs + (long long)group * Hz;
    actor_action_acts<G>(gm, gl, aw, nvl, a.cs, ox, oy, Hz, count, acts);
    const float total = actor_softmax<G>(gm, gl, Hz, acts);
    for (int hc0 = 0; hc0 < Hz; hc0 += G) {
        const int hc = hc0 + gl;
        const bool act = hc < Hz;
        float delta = 0.0f;
        if (act)
            delta = (mimic ? beta : (tdErrorAction > 0.0f ? beta : -beta)) *
                    ((hc == target ? 1.0f : 0.0f) - __ldcg(acts + hc) / fmaxf(0.0001f, total));
        for (int v = 0; v < nvl; ++v) {
            const ogn_wset ws = aw[v];
            const ogn_rf rf = *ws.rf;
            const int lx = __ldg(rf.lox + ox), nxv = __ldg(rf.nx + ox), ly = __ldg(rf.loy + oy), nyv = __ldg(rf.ny + oy);
            float *w = ws.wf + (col_block(rf, ox, oy) - ws.wf_base) + hc;
            delta_row<G>(gm, gl, act, w, a.cs[v], delta, lx, ly, nyv, nxv * nyv, rf.Vy, rf.Vz, rf.Hz);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight (re)layout. One block per hidden column; element e of the column block is addressed in the *device* order.
// which: 0 = F [slot][vz][hz], 1 = R [slot][hz][vz]; CSR order is [hz][slot][vz].
__device__ __forceinline__ long long csr_index_of(int which, long long e, int nslots, int Vz, int Hz) {
    int s, vz, hz;
    if (which == 0) { hz = (int)(e % Hz); vz = (int)((e / Hz) % Vz); s = (int)(e / ((long long)Hz * Vz)); }
    else { vz = (int)(e % Vz); hz = (int)((e / Vz) % Hz); s = (int)(e / ((long long)Hz * Vz)); }
    return ((long long)hz * nslots + s) * Vz + vz;
}

__global__ void k_layout(ogn_rf rf, float *dev, long long devBase, int which, int col0, float *csr, long long csrBase, int toCsr) {
    const int col = col0 + blockIdx.x;
    const int ox = col / rf.Hy, oy = col % rf.Hy;
    const int nslots = rf.nx[ox] * rf.ny[oy];
    const long long blk = col_block(rf, ox, oy);
    const long long n = (long long)nslots * rf.Vz * rf.Hz;
    float *d = dev + (blk - devBase);
    float *c = csr + (blk - csrBase);
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        const long long ci = csr_index_of(which, e, nslots, rf.Vz, rf.Hz);
        if (toCsr) c[ci] = d[e];
        else d[e] = c[ci];
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_init_hash(ogn_rf rf, float *wf, long long wfBase, int fx0, int fx1, float *wr, long long wrBase, int rx0, int rx1,
                            int col0, uint64_t key, float lo, float hi) {
    const int col = col0 + blockIdx.x;
    const int ox = col / rf.Hy, oy = col % rf.Hy;
    const int nslots = rf.nx[ox] * rf.ny[oy];
    const long long blk = col_block(rf, ox, oy);
    const long long n = (long long)nslots * rf.Vz * rf.Hz;
    const bool inF = wf && ox >= fx0 && ox < fx1, inR = wr && ox >= rx0 && ox < rx1;
    const float span = hi - lo;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) { // e in F order so the F stores coalesce
        const int hz = (int)(e % rf.Hz), vz = (int)((e / rf.Hz) % rf.Vz), s = (int)(e / ((long long)rf.Hz * rf.Vz));
        const long long ci = blk + ((long long)hz * nslots + s) * rf.Vz + vz; // global CSR index of this weight
        const uint32_t u24 = (uint32_t)(mix64(key + (uint64_t)ci * 0xD1B54A32D192ED03ull) >> 40);
        const float val = (float)u24 * (1.0f / 16777216.0f) * span + lo;
        if (inF) wf[blk - wfBase + e] = val;
        if (inR) wr[blk - wrBase + ((long long)s * rf.Hz + hz) * rf.Vz + vz] = val;
    }
}

__global__ void k_fr_mismatch(ogn_rf rf, const float *wf, long long wfBase, const float *wr, long long wrBase, int col0,
                              unsigned long long *count) {
    const int col = col0 + blockIdx.x;
    const int ox = col / rf.Hy, oy = col % rf.Hy;
    const int nslots = rf.nx[ox] * rf.ny[oy];
    const long long blk = col_block(rf, ox, oy);
    const long long n = (long long)nslots * rf.Vz * rf.Hz;
    unsigned long long bad = 0;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        const int hz = (int)(e % rf.Hz), vz = (int)((e / rf.Hz) % rf.Vz), s = (int)(e / ((long long)rf.Hz * rf.Vz));
        const float a = wf[blk - wfBase + e], b = wr[blk - wrBase + ((long long)s * rf.Hz + hz) * rf.Vz + vz];
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

inline int group_lanes(int z) {
    int g = 1;
    while (g < z && g < 32) g <<= 1;
    return g;
}
inline int launch_status() { return (int)cudaGetLastError(); }

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

int ogn_sc_forward(const ogn_wset *wsets, int nvl, const ogn_cs_args *cs, int Hx, int Hy, int Hz, int x0, int x1, int batch,
                   int *hidden_cs, int *hidden_cs2, void *stream) {
    if (nvl > OGN_MAX_VL) return (int)cudaErrorInvalidValue;
    const int ncols = (x1 - x0) * Hy;
    if (ncols <= 0 || batch <= 0) return 0;
    const int g = group_lanes(Hz);
    dim3 grid((unsigned)(((long long)ncols * g + kThreads - 1) / kThreads), (unsigned)batch);
    OGN_DISPATCH_G(g, (k_sc_forward<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(wsets, nvl, *cs, Hx, Hy, Hz, x0, ncols,
                                                                                 hidden_cs, hidden_cs2)));
    return launch_status();
}

int ogn_sc_learn(const ogn_wset *wsets, int v0, int nv, const ogn_cs_args *cs, const int *hidden_cs, const int *hidden_cs_prev,
                 int Hx, int Hy, int Hz, int Vx, int Vy, int Vz, int vx0, int vx1, float alpha, int batch,
                 unsigned long long *update_count, void *stream) {
    (void)Hx; (void)Hy; (void)Hz; (void)Vx;
    const int ncols = (vx1 - vx0) * Vy;
    if (ncols <= 0 || batch <= 0 || nv <= 0) return 0;
    const int g = group_lanes(Vz);
    dim3 grid((unsigned)(((long long)ncols * g + kThreads - 1) / kThreads), (unsigned)batch, (unsigned)nv);
    OGN_DISPATCH_G(g, (k_sc_learn<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(wsets, v0, *cs, hidden_cs, hidden_cs_prev, vx0,
                                                                               ncols, alpha, update_count)));
    return launch_status();
}

int ogn_pred_step(const ogn_wset *wsets, int nvl, const ogn_pred_args *args, int Hx, int Hy, int Hz, int x0, int x1, int batch,
                  int learn, int activate, const int *target_cs, float alpha, float *activations, int *hidden_cs, void *stream) {
    if (nvl > 8) return (int)cudaErrorInvalidValue;
    const int ncols = (x1 - x0) * Hy;
    if (ncols <= 0 || batch <= 0) return 0;
    const int g = group_lanes(Hz);
    dim3 grid((unsigned)(((long long)ncols * g + kThreads - 1) / kThreads), (unsigned)batch);
    OGN_DISPATCH_G(g, (k_pred_step<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(wsets, nvl, *args, Hx, Hy, Hz, x0, ncols, learn,
                                                                                activate, target_cs, alpha, activations,
                                                                                hidden_cs)));
    return launch_status();
}

int ogn_actor_forward(const ogn_wset *vw, const ogn_wset *aw, int nvl, const ogn_pred_args *args, int Hx, int Hy, int Hz,
                      const float *u01, float *hidden_values, float *hidden_activations, int *hidden_cs, void *stream) {
    if (nvl > 8) return (int)cudaErrorInvalidValue;
    const int g = group_lanes(Hz);
    dim3 grid((unsigned)(((long long)Hx * Hy * g + kThreads - 1) / kThreads));
    OGN_DISPATCH_G(g, (k_actor_forward<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(vw, aw, nvl, *args, Hx, Hy, Hz, u01,
                                                                                    hidden_values, hidden_activations,
                                                                                    hidden_cs)));
    return launch_status();
}

int ogn_actor_learn(const ogn_wset *vw, const ogn_wset *aw, int nvl, const ogn_pred_args *args, int Hx, int Hy, int Hz,
                    const int *target_cs_prev, const float *hidden_values_prev, const float *hidden_values, float q, float g_,
                    float alpha, float beta, int mimic, float *hidden_activations, void *stream) {
    if (nvl > 8) return (int)cudaErrorInvalidValue;
    const int g = group_lanes(Hz);
    dim3 grid((unsigned)(((long long)Hx * Hy * g + kThreads - 1) / kThreads));
    OGN_DISPATCH_G(g, (k_actor_learn<G><<<grid, kThreads, 0, (cudaStream_t)stream>>>(vw, aw, nvl, *args, Hx, Hy, Hz, target_cs_prev,
                                                                                  hidden_values_prev, hidden_values, q, g_,
                                                                                  alpha, beta, mimic, hidden_activations)));
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
