// ogn_vec.cuh — the 128-bit ("octet") kernels of the Sparse Predictive Hierarchy step: the fast path taken whenever a
// gathered row holds 8, 16 or 32 cells, i.e. whenever the packed F / R layouts of ogn_kernels_api.h put exactly one
// 128-byte line behind every (column pack, receptive-field slot, active input cell). Included by ogn_kernels.cu.
//
// Execution model: an OCTET of 8 lanes owns one 128-byte line; lane `sub` holds floats [4*sub, 4*sub+4) of it, which
// belong to column p = 4*sub / Z of the pack and are that column's cells (4*sub % Z) .. +3. One LDG.128 / STG.128 warp
// instruction moves four full 128-byte lines (512 B) and a batch of 8*NCH of them is issued before the first dependent
// add. The lanes of an octet take turns turning "their" slot of the next chunk of 8 slots into a line descriptor (one
// CSDR read, prefetched one chunk ahead); descriptors travel by width-8 shuffles. Every cell is accumulated in the
// reference's slot order (loads are batched, adds are sequential), so sums stay bit-identical to
// SparseMatrix::multiplyOHVs / multiplyOHVsT (SparseMatrix.cpp:260-294).
//
// Two work mappings (template parameter W = lanes per pack):
//   W = 8   one octet per pack, four independent packs per warp: the bandwidth mapping for launches that fill the machine.
//   W = 32  ("wide") one WARP per pack: the four octets take the pack's chunks round-robin, so a pack has four times
//           the lines in flight and a quarter of the dependent round trips; the running sum is handed from octet to
//           octet in chunk order (one owner adds at a time, the owner's sum is broadcast by shuffle), which keeps the
//           reference's summation order. This is the latency mapping: small layers (whole-hierarchy kernel, BASELINE
//           configs 1-3) and the 1/8-size launches of a sharded layer, where four times as many scheduling units
//           remove the wave-quantisation tail.
// CSDR access is a policy (CsNc: read-only path; CsCg: L2-coherent, for CSDRs produced earlier in the same launch;
// CsSm: a tile staged in shared memory by the bulk-copy engine, see "TMA staging" below).
#pragma once

namespace vec {

constexpr int kLineBits = 20;                    // descriptor: line index in the low 20 bits, union dy above it
constexpr unsigned kLineMask = (1u << kLineBits) - 1u;
constexpr unsigned kNoDy = 0xFFFu;               // "slot outside the pack's field" marker in the dy bits

__device__ __forceinline__ float4 ld_stream(const float *p) { // read-once weights: do not allocate in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_plain(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld_cg(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st_plain(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
template <bool kStream> __device__ __forceinline__ float4 ld_w(const float *p) { return kStream ? ld_stream(p) : ld_plain(p); }

struct CsNc { static __device__ __forceinline__ int ld(const int *p) { return __ldg(p); } };  // written by an earlier launch
struct CsCg { static __device__ __forceinline__ int ld(const int *p) { return __ldcg(p); } }; // written earlier in THIS launch
struct CsSm { static __device__ __forceinline__ int ld(const int *p) { return *p; } };        // shared-memory tile

__device__ __forceinline__ unsigned octet_mask() { return 0xffu << (threadIdx.x & 24); }
__device__ __forceinline__ float4 shfl4(unsigned m, const float4 &v, int src) {
    float4 r;
    r.x = __shfl_sync(m, v.x, src); r.y = __shfl_sync(m, v.y, src); r.z = __shfl_sync(m, v.z, src); r.w = __shfl_sync(m, v.w, src);
    return r;
}

// Walks the union field of one pack; `next()` returns this lane's descriptor of its slot of its next chunk:
// (slot * Vz + active cell) | dyU << 20 (dy bits = kNoDy for the clamped slots past the end of the field). A lane's slots
// are first, first + STEP, ... (STEP = 8: every chunk; STEP = 32: every fourth chunk, wide mapping). The CSDR is read at
// cs[(lx + dx) * rs + lyU + dy]: rs = Vy for the CSDR itself, the tile's row stride for a staged tile (cs is then the
// tile's virtual origin).
template <class CS, int STEP>
struct SlotWalker {
    const int *cs;
    int rs, lx, lyU, nyU, nxv, n, Vz;
    int s, dx, dy, cv;
    __device__ __forceinline__ void fetch() {
        const bool in = s < n;
        cv = CS::ld(cs + (lx + (in ? dx : nxv - 1)) * rs + lyU + (in ? dy : nyU - 1));
    }
    __device__ __forceinline__ void init(const int *cs_, int rs_, const PackCtx &c, int Vz_, int first) {
        cs = cs_; rs = rs_; lx = c.lx; lyU = c.lyU; nyU = c.nyU; n = c.nslotsU; nxv = n / max(nyU, 1); Vz = Vz_;
        s = first; dx = s / max(nyU, 1); dy = s - dx * nyU;
        fetch();
    }
    __device__ __forceinline__ unsigned next() {
        const bool in = s < n;
        const unsigned d = (unsigned)((in ? s : n - 1) * Vz + cv) | ((in ? (unsigned)dy : kNoDy) << kLineBits);
        s += STEP; dy += STEP;
        while (dy >= nyU) { dy -= nyU; ++dx; }
        fetch();
        return d;
    }
};

// Σ over the lane's column's field of its 4 cells of W_F[...][cs[visible column]] (multiplyOHVs, SparseMatrix.cpp:260-276).
// wl = pack block + 4*sub. W = 8: the octet walks every chunk, 8*NCH line loads per batch. W = 32: octet `oct` of the
// warp loads chunks oct, oct+4, ...; the adds run in chunk order with the sum handed from owner to owner.
template <bool kStream, int NCH, int W, class CS>
__device__ __forceinline__ float4 gather(unsigned om, int sub, int oct, const float *wl, const int *cs, int rs, const PackCtx &c, int Vz) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c.nslotsU <= 0) return acc;
    if constexpr (W == 8) {
        SlotWalker<CS, 8> sw;
        sw.init(cs, rs, c, Vz, sub);
        for (int ch = 0; ch < c.nslotsU; ch += 8 * NCH) {
            unsigned d[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) d[k] = sw.next();
            float4 w[8 * NCH];
            unsigned dj[8 * NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    dj[k * 8 + j] = __shfl_sync(om, d[k], j, 8);
                    w[k * 8 + j] = ld_w<kStream>(wl + ((size_t)(dj[k * 8 + j] & kLineMask) << 5));
                }
#pragma unroll
            for (int j = 0; j < 8 * NCH; ++j)
                if ((unsigned)((int)(dj[j] >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                    acc.x += w[j].x; acc.y += w[j].y; acc.z += w[j].z; acc.w += w[j].w;
                }
        }
    } else {
        SlotWalker<CS, 32> sw;
        sw.init(cs, rs, c, Vz, oct * 8 + sub);
        const int nch = (c.nslotsU + 7) >> 3;
        for (int ch0 = 0; ch0 < nch; ch0 += 4 * NCH) { // one round: chunks ch0 + 4k + o, k < NCH, o < 4
            unsigned d[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) d[k] = sw.next();
            float4 w[8 * NCH];
            unsigned dj[8 * NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    dj[k * 8 + j] = __shfl_sync(om, d[k], j, 8);
                    w[k * 8 + j] = ld_w<kStream>(wl + ((size_t)(dj[k * 8 + j] & kLineMask) << 5));
                }
#pragma unroll
            for (int k = 0; k < NCH; ++k)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (ch0 + 4 * k + o < nch) { // warp-uniform
                        if (oct == o) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if ((unsigned)((int)(dj[k * 8 + j] >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                                    acc.x += w[k * 8 + j].x; acc.y += w[k * 8 + j].y; acc.z += w[k * 8 + j].z; acc.w += w[k * 8 + j].w;
                                }
                        }
                        acc = shfl4(0xffffffffu, acc, o * 8 + sub); // the owner's running sum goes to every octet
                    }
                }
        }
    }
    return acc;
}

// W_F[...][cs[visible column]] += delta over the lane's column's field (deltaOHVs, SparseMatrix.cpp:488-501). No ordering
// between slots: in the wide mapping the octets simply split the chunks.
template <int NCH, int W, class CS>
__device__ __forceinline__ void add_delta(unsigned om, int sub, int oct, float *wl, const int *cs, int rs, const float4 &delta, const PackCtx &c,
                                          int Vz) {
    if (c.nslotsU <= 0) return;
    SlotWalker<CS, W> sw;
    sw.init(cs, rs, c, Vz, W == 8 ? sub : oct * 8 + sub);
    const int nch = (c.nslotsU + 7) >> 3;
    constexpr int NO = W / 8;
    for (int ch0 = 0; ch0 < nch; ch0 += NO * NCH) {
        unsigned d[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) d[k] = sw.next();
        float4 w[8 * NCH];
        unsigned dj[8 * NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dj[k * 8 + j] = __shfl_sync(om, d[k], j, 8);
                w[k * 8 + j] = ld_plain(wl + ((size_t)(dj[k * 8 + j] & kLineMask) << 5));
            }
#pragma unroll
        for (int j = 0; j < 8 * NCH; ++j)
            if ((unsigned)((int)(dj[j] >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                float4 v = w[j];
                v.x += delta.x; v.y += delta.y; v.z += delta.z; v.w += delta.w;
                st_plain(wl + ((size_t)(dj[j] & kLineMask) << 5), v);
            }
    }
}

// ---- rows staged in shared memory (octet mapping) ----------------------------------------------------------------------
// The same walks as gather<., NS, 8> / add_delta<NS, 8>, but the 8 * NS row pieces a lane has in flight are 16-byte cp.async
// copies into the lane's own column of a shared-memory stage (stage[r * blockDim.x] for r < 8 * NS) instead of registers:
// rows in flight cost shared memory, so one round can hold 24 rows per lane (96 KB per block of 256, two blocks per SM)
// where the register version holds 8. tools/microbench_patterns.cu: for this access pattern (one 128-byte row out of Vz per
// slot) 27 staged rows per lane reach 5.39 TB/s read-modify-write + gather at two blocks per SM, 9 rows in registers 3.99.
// A lane only ever reads the pieces it copied itself, so cp.async.wait_group is the only synchronisation needed.
__device__ __forceinline__ void cp_async16(float4 *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int NS, class CS>
__device__ __forceinline__ float4 gather_staged(unsigned om, int sub, const float *wl, const int *cs, int rs, const PackCtx &c, int Vz, float4 *stage) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c.nslotsU <= 0) return acc;
    const int stride = blockDim.x;
    SlotWalker<CS, 8> sw;
    sw.init(cs, rs, c, Vz, sub);
    for (int ch = 0; ch < c.nslotsU; ch += 8 * NS) {
        unsigned d[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) d[k] = sw.next();
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned dj = __shfl_sync(om, d[k], j, 8);
                cp_async16(stage + (k * 8 + j) * stride, wl + ((size_t)(dj & kLineMask) << 5));
            }
        cp_async_wait_all();
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned dj = __shfl_sync(om, d[k], j, 8);
                if ((unsigned)((int)(dj >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                    const float4 w = stage[(k * 8 + j) * stride];
                    acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
                }
            }
    }
    return acc;
}

template <int NS, class CS>
__device__ __forceinline__ void add_delta_staged(unsigned om, int sub, float *wl, const int *cs, int rs, const float4 &delta, const PackCtx &c, int Vz,
                                                 float4 *stage) {
    if (c.nslotsU <= 0) return;
    const int stride = blockDim.x;
    SlotWalker<CS, 8> sw;
    sw.init(cs, rs, c, Vz, sub);
    for (int ch = 0; ch < c.nslotsU; ch += 8 * NS) {
        unsigned d[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) d[k] = sw.next();
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned dj = __shfl_sync(om, d[k], j, 8);
                cp_async16(stage + (k * 8 + j) * stride, wl + ((size_t)(dj & kLineMask) << 5));
            }
        cp_async_wait_all();
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned dj = __shfl_sync(om, d[k], j, 8);
                if ((unsigned)((int)(dj >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                    float4 v = stage[(k * 8 + j) * stride];
                    v.x += delta.x; v.y += delta.y; v.z += delta.z; v.w += delta.w;
                    st_plain(wl + ((size_t)(dj & kLineMask) << 5), v);
                }
            }
    }
}

// First-max arg-max of the reference's forward scans (`if (sum > max)` from (0, -999999), SparseCoder.cpp:31-38,
// Predictor.cpp:36-43) over the Z cells of a column spread over LC = Z/4 lanes x 4 cells. NaN never wins.
template <int Z>
__device__ __forceinline__ int column_argmax_fwd(unsigned om, int hz0, const float4 &s) {
    const float a[4] = {s.x, s.y, s.z, s.w};
    float v = -999999.0f;
    int i = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x = a[k] > -999999.0f ? a[k] : -999999.0f;
        if (k == 0 || x > v) { v = x; i = hz0 + k; }
    }
#pragma unroll
    for (int o = Z / 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(om, v, o, 8);
        const int oi = __shfl_xor_sync(om, i, o, 8);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    return v > -999999.0f ? i : 0;
}

struct PackId { int ox, q, p, oy, hz0; bool colValid; };
template <int Z> __device__ __forceinline__ PackId pack_id(int pack, int packsPerRow, int x0, int Hy, int sub) {
    PackId id;
    id.ox = x0 + pack / packsPerRow; id.q = pack % packsPerRow;
    id.p = (sub * 4) / Z; id.hz0 = (sub * 4) % Z;
    id.oy = id.q * (32 / Z) + id.p;
    id.colValid = id.oy < Hy;
    return id;
}

// A CSDR as the walkers see it: base pointer (of this member) and row stride. tile == nullptr: the CSDR in global memory.
struct CsView { const int *p; int rs; };

// ---------------------------------------------------------------------------------------------------------------
// K2 for one pack: SparseCoder::forward (SparseCoder.cpp:13-43), Hz = Z in {8, 16, 32}. `tiles` (optional) holds one
// staged CSDR view per visible layer.
template <int Z, int NCH, int W, class CS>
__device__ __forceinline__ void sc_forward_pack(const ogn_fwd_args &A, int Hx, int Hy, int x0, int pack, int packsPerRow, int member,
                                                int *out0, int *out1, const CsView *tiles) {
    const int sub = threadIdx.x & 7, oct = (threadIdx.x >> 3) & 3;
    const unsigned om = octet_mask();
    const PackId id = pack_id<Z>(pack, packsPerRow, x0, Hy, sub);
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (W == 32 && A.nvl >= 3) {
        // wide mapping with several visible layers (temporalHorizon x inputs): the reference sums every visible layer into
        // its own accumulator and adds the results in layer order (SparseCoder.cpp:23-40), so the four octets of the warp
        // each take ONE visible layer (a whole octet-per-pack gather, 16 lines in flight per lane) and the per-layer sums
        // are combined in order afterwards: the layers' dependent chains (tables, CSDR, weights) run side by side
        // instead of one after the other.
        for (int base = 0; base < A.nvl; base += 4) {
            const int v = base + oct;
            float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (v < A.nvl) {
                const ogn_fwd_vl &vl = A.vl[v];
                const ogn_rf &rf = A.geo[vl.geo];
                const PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
                const float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
                CsView cv;
                if (tiles) cv = tiles[v];
                else { cv.p = vl.cs + (long long)member * rf.Vx * rf.Vy; cv.rs = rf.Vy; }
                g = gather<true, 2, 8, CS>(om, sub, 0, wl, cv.p, cv.rs, c, rf.Vz);
            }
            __syncwarp();
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (base + o < A.nvl) {
                    const float4 t = shfl4(0xffffffffu, g, o * 8 + sub);
                    sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
                }
        }
    } else {
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_fwd_vl &vl = A.vl[v];
            const ogn_rf &rf = A.geo[vl.geo];
            const PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
            const float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
            CsView cv;
            if (tiles) cv = tiles[v];
            else { cv.p = vl.cs + (long long)member * rf.Vx * rf.Vy; cv.rs = rf.Vy; }
            const float4 g = gather<true, NCH, W, CS>(om, sub, oct, wl, cv.p, cv.rs, c, rf.Vz);
            sum.x += g.x; sum.y += g.y; sum.z += g.z; sum.w += g.w;
        }
    }
    const int best = column_argmax_fwd<Z>(om, id.hz0, sum);
    if (id.hz0 == 0 && id.colValid && (W == 8 || oct == 0)) {
        const long long o = (long long)member * Hx * Hy + id.ox * Hy + id.oy;
        out0[o] = best;
        if (out1) out1[o] = best;
    }
}

// K6+K7 for one pack: Predictor::learn (Predictor.cpp:49-70) then Predictor::forward (:13-47), Hz = Z in {8, 16, 32}.
// tiles (optional): [v] = staged view of vl[v].cs, [OGN_MAX_PRED_VL + v] = staged view of vl[v].cs_prev.
template <int Z, int NCH, int W, class CS, int STG = 0>
__device__ __forceinline__ void pred_step_pack(const ogn_pred_args &A, int Hx, int Hy, int x0, int pack, int packsPerRow, int member, int learn,
                                               int activate, const int *target, float alpha, float *acts, int *hiddenCs,
                                               const CsView *tiles, float4 *stage = nullptr) {
    const int sub = threadIdx.x & 7, oct = (threadIdx.x >> 3) & 3;
    const unsigned om = octet_mask();
    const PackId id = pack_id<Z>(pack, packsPerRow, x0, Hy, sub);
    const long long col = (long long)member * Hx * Hy + id.ox * Hy + (id.colValid ? id.oy : Hy - 1);
    float *act4 = acts + col * Z + id.hz0;

    if (learn) {
        const int t = CS::ld(target + col); // a CsSm instantiation never stages the target: it is read from global memory
        const float4 a = ld_plain(act4);
        float4 delta;
        delta.x = alpha * ((id.hz0 + 0 == t ? 1.0f : -1.0f) - ogn_tanhf(a.x));
        delta.y = alpha * ((id.hz0 + 1 == t ? 1.0f : -1.0f) - ogn_tanhf(a.y));
        delta.z = alpha * ((id.hz0 + 2 == t ? 1.0f : -1.0f) - ogn_tanhf(a.z));
        delta.w = alpha * ((id.hz0 + 3 == t ? 1.0f : -1.0f) - ogn_tanhf(a.w));
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_pred_vl &vl = A.vl[v];
            const ogn_rf &rf = A.geo[vl.geo];
            PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
            if (!id.colValid) c.dn = 0;
            float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
            CsView cv;
            if (tiles) cv = tiles[OGN_MAX_PRED_VL + v];
            else { cv.p = vl.cs_prev + (long long)member * rf.Vx * rf.Vy; cv.rs = rf.Vy; }
            if constexpr (STG > 0) add_delta_staged<STG, CS>(om, sub, wl, cv.p, cv.rs, delta, c, rf.Vz, stage);
            else add_delta<NCH, W, CS>(om, sub, oct, wl, cv.p, cv.rs, delta, c, rf.Vz);
        }
        if (W == 32) __syncwarp(); // the activate pass re-reads lines another octet of this warp just stored
    }
    if (!activate) return;

    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int count = 0;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_pred_vl &vl = A.vl[v];
        const ogn_rf &rf = A.geo[vl.geo];
        const PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
        const float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
        CsView cv;
        if (tiles) cv = tiles[v];
        else { cv.p = vl.cs + (long long)member * rf.Vx * rf.Vy; cv.rs = rf.Vy; }
        float4 g;
        if constexpr (STG > 0) g = gather_staged<STG, CS>(om, sub, wl, cv.p, cv.rs, c, rf.Vz, stage);
        else g = gather<false, NCH, W, CS>(om, sub, oct, wl, cv.p, cv.rs, c, rf.Vz);
        sum.x += g.x; sum.y += g.y; sum.z += g.z; sum.w += g.w;
        count += c.dn * (c.nslotsU / max(c.nyU, 1)); // own field: nx * ny
    }
    const float cnt = (float)count;
    sum.x = div_ref(sum.x, cnt); sum.y = div_ref(sum.y, cnt); sum.z = div_ref(sum.z, cnt); sum.w = div_ref(sum.w, cnt);
    const bool writer = W == 8 || oct == 0;
    if (id.colValid && writer) st_plain(act4, sum);
    const int best = column_argmax_fwd<Z>(om, id.hz0, sum);
    if (id.hz0 == 0 && id.colValid && writer) hiddenCs[col] = best;
}

// K8: the grid copies the inputs into the *other* inputCsPrev buffer (never read by the launch that writes it)
template <class CS>
__device__ __forceinline__ void pred_copy_inputs(const ogn_pred_args &A, int member, int tid, int nth) {
    for (int v = 0; v < A.nvl; ++v) {
        if (!A.vl[v].cs_prev_out) continue;
        const int n = A.geo[A.vl[v].geo].Vx * A.geo[A.vl[v].geo].Vy;
        const long long mo = (long long)member * n;
        for (int i = tid; i < n; i += nth) A.vl[v].cs_prev_out[mo + i] = CS::ld(A.vl[v].cs + mo + i);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3, part 1 for one pack: the reconstruction of SparseCoder::learn (SparseCoder.cpp:45-70, multiplyOHVsT
// SparseMatrix.cpp:278-294, countT :215-221), Vz = Z in {8, 16, 32}. One pack = 32/Z y-adjacent visible columns; lane = 4
// visible cells. Visible columns whose reconstruction arg-max misses the target are appended to a work list for part 2.
template <class CS, int STEP>
struct CoverWalker {
    const int *hcs;
    int hx0, hy0U, ncyU, n, vx, vyBase;
    int k, ix, iy; // this lane's item of the chunk being described: hidden column (hx0 + ix, hy0U + iy)
    template <int Z> __device__ __forceinline__ unsigned long long describe(const ogn_rf &rf) const {
        constexpr int PR = 32 / Z;
        const bool in = k < n;
        const int ox = hx0 + (in ? ix : (n - 1) / ncyU), oy = hy0U + (in ? iy : (n - 1) % ncyU);
        const int hc = CS::ld(hcs + ox * rf.Hy + oy);
        const int ylo = __ldg(rf.loy + oy), yn = __ldg(rf.ny + oy);
        const int dx = vx - __ldg(rf.lox + ox);
        const int gi = vyBase / PR - ylo / PR; // visible-y group index inside this hidden column's field
        unsigned long long d =
            (unsigned long long)(r_block(rf, ox, oy) + (((long long)dx * __ldg(rf.ngr + oy) + gi) * rf.Hz + hc) * 32);
        if (in) {
#pragma unroll
            for (int pv = 0; pv < PR; ++pv)
                if ((unsigned)(vyBase + pv - ylo) < (unsigned)yn && vyBase + pv < rf.Vy) d |= 1ull << pv;
        }
        return d;
    }
    __device__ __forceinline__ void advance() {
        k += STEP; iy += STEP;
        while (iy >= ncyU) { iy -= ncyU; ++ix; }
    }
};

template <int Z, int NCH, int W, class CS>
__device__ __forceinline__ void sc_recon_pack(const ogn_learn_args &A, int vli, const int *hcsAll, int vx0, int pack, int packsPerRow,
                                              int member, int listTag, int2 *work, unsigned *workCount) {
    constexpr int PR = 32 / Z;
    const int sub = threadIdx.x & 7, oct = (threadIdx.x >> 3) & 3;
    const unsigned om = octet_mask();
    const ogn_learn_vl &vl = A.vl[vli];
    const ogn_rf &rf = A.geo;
    const int vx = vx0 + pack / packsPerRow, vyBase = (pack % packsPerRow) * PR;
    const int pv = (sub * 4) / Z, vc0 = (sub * 4) % Z;
    const int vy = vyBase + pv;
    const bool colValid = vy < rf.Vy;
    const int vyc = colValid ? vy : rf.Vy - 1;
    const long long vcol = (long long)member * rf.Vx * rf.Vy + vx * rf.Vy + vyc;
    const float *wr = vl.wr + (long long)member * vl.r_mstride + sub * 4;

    // cover sets: own (for the count) and the pack's union along y (walked together)
    const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx);
    const int hy0 = __ldg(rf.rloy + vyc), hy1 = __ldg(rf.rhiy + vyc);
    const int ncx = max(0, hx1 - hx0 + 1), ncy = max(0, hy1 - hy0 + 1);
    const int ncov = ncx * ncy; // == countT(visibleIndex) / hiddenSize.z
    int hy0U = rf.Hy, hy1U = -1;
#pragma unroll
    for (int k = 0; k < PR; ++k) {
        const int y = vyBase + k;
        if (y < rf.Vy) {
            const int a = __ldg(rf.rloy + y), b = __ldg(rf.rhiy + y);
            if (b >= a) { hy0U = min(hy0U, a); hy1U = max(hy1U, b); }
        }
    }
    const int ncyU = max(0, hy1U - hy0U + 1), ncovU = ncx * ncyU;

    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (ncovU > 0) {
        CoverWalker<CS, W> cw;
        cw.hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
        cw.hx0 = hx0; cw.hy0U = hy0U; cw.ncyU = ncyU; cw.n = ncovU; cw.vx = vx; cw.vyBase = vyBase;
        cw.k = W == 8 ? sub : oct * 8 + sub; cw.ix = cw.k / ncyU; cw.iy = cw.k - cw.ix * ncyU;
        constexpr int NO = W / 8;
        const int nch = (ncovU + 7) >> 3;
        unsigned long long d[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) { d[k] = cw.template describe<Z>(rf); cw.advance(); }
        for (int ch0 = 0; ch0 < nch; ch0 += NO * NCH) {
            float4 w[8 * NCH];
            unsigned vm = 0;
#pragma unroll
            for (int k = 0; k < NCH; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned long long dj = __shfl_sync(om, d[k], j, 8);
                    vm |= (unsigned)((dj >> pv) & 1ull) << (k * 8 + j);
                    w[k * 8 + j] = ld_stream(wr + (dj & ~31ull));
                }
            // descriptors of the next batch: their table / CSDR reads overlap the weight loads in flight
#pragma unroll
            for (int k = 0; k < NCH; ++k) { d[k] = cw.template describe<Z>(rf); cw.advance(); }
            if constexpr (W == 8) {
#pragma unroll
                for (int j = 0; j < 8 * NCH; ++j)
                    if ((vm >> j) & 1u) { sum.x += w[j].x; sum.y += w[j].y; sum.z += w[j].z; sum.w += w[j].w; }
            } else {
#pragma unroll
                for (int k = 0; k < NCH; ++k)
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (ch0 + 4 * k + o < nch) {
                            if (oct == o) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if ((vm >> (k * 8 + j)) & 1u) {
                                        sum.x += w[k * 8 + j].x; sum.y += w[k * 8 + j].y; sum.z += w[k * 8 + j].z; sum.w += w[k * 8 + j].w;
                                    }
                            }
                            sum = shfl4(0xffffffffu, sum, o * 8 + sub);
                        }
                    }
            }
        }
    }
    const float cnt = (float)ncov;
    sum.x = div_ref(sum.x, cnt); sum.y = div_ref(sum.y, cnt); sum.z = div_ref(sum.z, cnt); sum.w = div_ref(sum.w, cnt);
    const bool writer = W == 8 || oct == 0;
    if (colValid && writer) st_plain(vl.recon + vcol * Z + vc0, sum);

    // arg-max with the reference's `sum > max || maxIndex == -1` rule (SparseCoder.cpp:60-66): element 0 is always taken
    // first, NaN elsewhere never wins
    const float a[4] = {sum.x, sum.y, sum.z, sum.w};
    float v = 0.0f;
    int i = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = a[k];
        if (vc0 + k == 0) x = (x != x) ? INFINITY : x;
        else x = (x != x) ? -INFINITY : x;
        if (k == 0 || x > v) { v = x; i = vc0 + k; }
    }
#pragma unroll
    for (int o = Z / 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(om, v, o, 8);
        const int oi = __shfl_xor_sync(om, i, o, 8);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (colValid && writer && vc0 == 0 && i != CS::ld(vl.cs + vcol)) {
        const unsigned slot = atomicAdd(workCount, 1u);
        work[slot] = make_int2(listTag, vx * rf.Vy + vy);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3, part 2: the delta rule on the weights of hidden columns whose state changed (deltaChangedOHVsT,
// SparseMatrix.cpp:572-590; SparseCoder.cpp:72-81) for the visible columns on the work list.
//
// Version 1 (any Hz): a work unit is one chunk of 32/LC * UN cover items of one listed column, one warp per unit; LC = Z/4
// lanes per (hidden column, visible column) pair: the R copy is rewritten with one 128-bit read-modify-write, the F copy
// with four scalar stores (one per forward row): 16 PARTIALLY written 32-byte sectors per pair, each of which the L2 has
// to fill from DRAM before it can write it back (ncu, round 1: 1.45 KB of DRAM traffic per pair, 2 row activations per
// sector: the kernel is bound by DRAM row activations, not bytes).
template <int Z, class CS, bool kCg>
__device__ __forceinline__ int sc_update_unit_v1(const ogn_learn_args &A, const int *hcsAll, const int *hcsPrevAll, const int2 wk, int k0,
                                                 float alpha) {
    constexpr int LC = Z / 4, IPW = 32 / LC, PR = 32 / Z, UN = 4;
    const ogn_rf &rf = A.geo;
    const int lane = threadIdx.x & 31;
    const int it = lane / LC, vc0 = (lane % LC) * 4;
    const long long fstep = (long long)rf.pf * rf.Hz; // floats between the forward rows of consecutive visible cells
    const int vx = wk.y / rf.Vy, vy = wk.y - vx * rf.Vy;
    const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx);
    const int hy0 = __ldg(rf.rloy + vy), hy1 = __ldg(rf.rhiy + vy);
    const int ncy = max(0, hy1 - hy0 + 1), ncov = max(0, hx1 - hx0 + 1) * ncy;
    if (k0 >= ncov) return 0;
    const int member = wk.x / A.nvl;
    const ogn_learn_vl &vl = A.vl[wk.x % A.nvl];
    const long long vcol = (long long)member * rf.Vx * rf.Vy + wk.y;
    const int *hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
    const int *hcsPrev = hcsPrevAll + (long long)member * rf.Hx * rf.Hy;
    float *wr = vl.wr + (long long)member * vl.r_mstride + vc0;
    float *wf = vl.wf + (long long)member * vl.f_mstride;

    long long offR[UN], offF[UN];
    int fl[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const int k = k0 + u * IPW + it;
        fl[u] = 0;
        offR[u] = 0; offF[u] = 0;
        if (k < ncov) {
            const int ox = hx0 + k / ncy, oy = hy0 + k % ncy;
            const int h = ox * rf.Hy + oy;
            const int hc = CS::ld(hcs + h);
            if (hc != CS::ld(hcsPrev + h)) {
                const int ylo = __ldg(rf.loy + oy);
                const int dx = vx - __ldg(rf.lox + ox);
                offR[u] = r_block(rf, ox, oy) +
                          ((((long long)dx * __ldg(rf.ngr + oy) + (vy / PR - ylo / PR)) * rf.Hz + hc) * PR + (vy % PR)) * Z;
                const int qh = oy / rf.pf, ph = oy % rf.pf;
                offF[u] = f_block(rf, ox, qh) +
                          (((long long)dx * __ldg(rf.nyf + qh) + (vy - __ldg(rf.loyf + qh))) * Z * rf.pf + ph) * rf.Hz + hc +
                          (long long)vc0 * fstep;
                fl[u] = 1 | ((ox >= A.fx0 && ox < A.fx1) ? 2 : 0);
            }
        }
    }
    float4 w[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
        if (fl[u]) w[u] = ld_plain(wr + offR[u]);
    // delta of this lane's 4 visible cells (SparseCoder.cpp:78) while the loads are in flight
    const int target = CS::ld(vl.cs + vcol);
    const float4 r = kCg ? ld_cg(vl.recon + vcol * Z + vc0) : ld_plain(vl.recon + vcol * Z + vc0);
    float4 delta;
    delta.x = alpha * ((vc0 + 0 == target ? 1.0f : 0.0f) - ogn_expf(r.x));
    delta.y = alpha * ((vc0 + 1 == target ? 1.0f : 0.0f) - ogn_expf(r.y));
    delta.z = alpha * ((vc0 + 2 == target ? 1.0f : 0.0f) - ogn_expf(r.z));
    delta.w = alpha * ((vc0 + 3 == target ? 1.0f : 0.0f) - ogn_expf(r.w));
    int nupd = 0;
#pragma unroll
    for (int u = 0; u < UN; ++u)
        if (fl[u]) {
            float4 nw = w[u];
            nw.x += delta.x; nw.y += delta.y; nw.z += delta.z; nw.w += delta.w;
            st_plain(wr + offR[u], nw);
            if (fl[u] & 2) {
                float *f = wf + offF[u];
                f[0] = nw.x; f[fstep] = nw.y; f[2 * fstep] = nw.z; f[3 * fstep] = nw.w;
            }
            if (vc0 == 0) nupd++;
        }
    return nupd;
}

// Version 2 (Hz a multiple of 8): FULL-sector writes into F. A pair's update changes one float (hidden cell hc) in each
// of the Vz forward rows of its slot; the 7 other floats of that 32-byte sector are the cells hc' of the same 8-aligned
// group, and their values sit, together with the row being rewritten, in 8 consecutive rows of the R copy (8 x Vz floats,
// contiguous up to the visible-y packing). So: SL = 2*Vz lanes load the 8 R rows of the group (one float4 each: lane =
// (hidden cell of the group, 4 visible cells)), the lanes holding row hc add the delta and store the R row back, and every
// lane stores its 4 values into F with 4 scalar stores whose 8 lanes per forward row cover exactly one 32-byte sector:
// no sector is partially written, so the L2 never reads F (no fill, half the row activations), and the F writes carry
// the group's unchanged cells, which are bit-identical in both copies (invariant checked by ogn_weights_fr_mismatch).
template <int Z, int GRP, class CS, bool kCg>
__device__ __forceinline__ int sc_update_unit_v2(const ogn_learn_args &A, const int *hcsAll, const int *hcsPrevAll, const int2 wk, int k0,
                                                 int batches, float alpha) {
    constexpr int LC = Z / 4, SL = GRP * LC, PPI = SL >= 32 ? 1 : 32 / SL, J = SL >= 32 ? SL / 32 : 1, PR = 32 / Z, UN = GRP == 8 ? 4 : 2;
    const ogn_rf &rf = A.geo;
    const int lane = threadIdx.x & 31;
    const long long fstep = (long long)rf.pf * rf.Hz;
    const int vx = wk.y / rf.Vy, vy = wk.y - vx * rf.Vy;
    const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx);
    const int hy0 = __ldg(rf.rloy + vy), hy1 = __ldg(rf.rhiy + vy);
    const int ncy = max(0, hy1 - hy0 + 1), ncov = max(0, hx1 - hx0 + 1) * ncy;
    if (k0 >= ncov) return 0;
    const int member = wk.x / A.nvl;
    const ogn_learn_vl &vl = A.vl[wk.x % A.nvl];
    const long long vcol = (long long)member * rf.Vx * rf.Vy + wk.y;
    const int *hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
    const int *hcsPrev = hcsPrevAll + (long long)member * rf.Hx * rf.Hy;
    float *wrm = vl.wr + (long long)member * vl.r_mstride;
    float *wfm = vl.wf + (long long)member * vl.f_mstride;
    const int pi = PPI > 1 ? lane / SL : 0; // which of the PPI pairs of an iteration this lane works on

    // slice j of this lane: hidden cell hz8[j] of the group, visible cells 4*q[j] .. +3
    int hz8[J], q4[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int sl = (PPI > 1 ? lane % SL : lane) + 32 * j;
        hz8[j] = sl % GRP; q4[j] = (sl / GRP) * 4; // the GRP lanes of one forward row are neighbours: one store instruction, one full granule
    }
    // delta of the visible cells of this lane's slices (SparseCoder.cpp:78): once per unit, shared by all its pairs
    const int target = CS::ld(vl.cs + vcol);
    float4 delta[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 r = kCg ? ld_cg(vl.recon + vcol * Z + q4[j]) : ld_plain(vl.recon + vcol * Z + q4[j]);
        delta[j].x = alpha * ((q4[j] + 0 == target ? 1.0f : 0.0f) - ogn_expf(r.x));
        delta[j].y = alpha * ((q4[j] + 1 == target ? 1.0f : 0.0f) - ogn_expf(r.y));
        delta[j].z = alpha * ((q4[j] + 2 == target ? 1.0f : 0.0f) - ogn_expf(r.z));
        delta[j].w = alpha * ((q4[j] + 3 == target ? 1.0f : 0.0f) - ogn_expf(r.w));
    }
    int nupd = 0;
    for (int b = 0; b < batches; ++b) {
        const int kb = k0 + b * (UN * PPI);
        if (kb >= ncov) break;
        long long offR[UN], offF[UN];
        int fl[UN], hcl[UN];
        float4 w[UN][J];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int k = kb + u * PPI + pi;
            fl[u] = 0; hcl[u] = 0; offR[u] = 0; offF[u] = 0;
            if (k < ncov) {
                const int ox = hx0 + k / ncy, oy = hy0 + k % ncy;
                const int h = ox * rf.Hy + oy;
                const int hc = CS::ld(hcs + h);
                if (hc != CS::ld(hcsPrev + h)) {
                    const int ylo = __ldg(rf.loy + oy);
                    const int dx = vx - __ldg(rf.lox + ox);
                    const int g8 = hc & ~(GRP - 1);
                    // R row of hidden cell g8 (the group's first), this visible column; consecutive hidden cells: + PR * Z floats
                    offR[u] = r_block(rf, ox, oy) +
                              ((((long long)dx * __ldg(rf.ngr + oy) + (vy / PR - ylo / PR)) * rf.Hz + g8) * PR + (vy % PR)) * Z;
                    const int qh = oy / rf.pf, ph = oy % rf.pf;
                    // F: forward row of visible cell 0, hidden cell g8; consecutive visible cells: + fstep floats
                    offF[u] = f_block(rf, ox, qh) +
                              (((long long)dx * __ldg(rf.nyf + qh) + (vy - __ldg(rf.loyf + qh))) * Z * rf.pf + ph) * rf.Hz + g8;
                    fl[u] = 1 | ((ox >= A.fx0 && ox < A.fx1) ? 2 : 0);
                    hcl[u] = hc & (GRP - 1);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (fl[u]) w[u][j] = ld_plain(wrm + offR[u] + (long long)hz8[j] * (PR * Z) + q4[j]);
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (fl[u]) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float4 nw = w[u][j];
                    if (hz8[j] == hcl[u]) {
                        nw.x += delta[j].x; nw.y += delta[j].y; nw.z += delta[j].z; nw.w += delta[j].w;
                        st_plain(wrm + offR[u] + (long long)hz8[j] * (PR * Z) + q4[j], nw);
                        if (q4[j] == 0) nupd++;
                    }
                    if (fl[u] & 2) {
                        float *f = wfm + offF[u] + (long long)q4[j] * fstep + hz8[j];
                        f[0] = nw.x; f[fstep] = nw.y; f[2 * fstep] = nw.z; f[3 * fstep] = nw.w;
                    }
                }
            }
    }
    return nupd;
}
template <int Z, int GRP = 8> struct UpdV2 {
    static constexpr int SL = GRP * Z / 4, PPI = SL >= 32 ? 1 : 32 / SL, ITEMS = PPI * (GRP == 8 ? 4 : 2); // cover items per batch; a unit is `batches` of them
};
template <int Z> struct UpdV1 {
    static constexpr int ITEMS = (32 / (Z / 4)) * 4;
};

// ---------------------------------------------------------------------------------------------------------------
// TMA staging of the input CSDRs (north_star: "input CSDRs are staged in shared memory with TMA"). A block of the W = 8
// kernels works on vt/8 y-consecutive packs of ONE hidden row, whose fields cover a window of at most max_nx visible rows
// x (a few dozen) visible columns of every visible layer. One elected thread queues one bulk copy (cp.async.bulk, the
// TMA engine's linear mode: SASS UBLKCP) per window row into shared memory and arms an mbarrier with the byte count;
// everybody waits on the barrier's phase and then builds line descriptors from shared memory (LDS, 29 cycles) instead
// of L2 (250+ cycles), with no CSDR prefetch registers. Rows start and end on 16-byte boundaries (Vy % 4 == 0 is the
// host-side eligibility test).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_init_nofence(unsigned long long *bar) { // caller fences once after the last init
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    for (int it = 0; it < (1 << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap(); // a copy that never lands is a bug: fail the launch instead of hanging the device
}

// window of the block's packs in one geometry: visible rows [lx, lx + nxv), visible columns [y0a, y0a + wa), wa % 4 == 0
struct TileWin { int lx, nxv, y0a, wa; };
__device__ __forceinline__ TileWin tile_window(const ogn_rf &rf, int ox, int q0, int q1) {
    TileWin t;
    t.lx = __ldg(rf.lox + ox); t.nxv = __ldg(rf.nx + ox);
    const int y0 = __ldg(rf.loyf + q0), y1 = __ldg(rf.loyf + q1) + __ldg(rf.nyf + q1);
    t.y0a = y0 & ~3;
    t.wa = ((y1 + 3) & ~3) - t.y0a;
    return t;
}
// queue the copies of one window (elected thread); returns the bytes it will deliver
__device__ __forceinline__ unsigned tile_issue(int *dst, const int *csMember, int Vy, const TileWin &t, unsigned long long *bar) {
    for (int r = 0; r < t.nxv; ++r) bulk_g2s(dst + r * t.wa, csMember + (long long)(t.lx + r) * Vy + t.y0a, (unsigned)t.wa * 4u, bar);
    return (unsigned)(t.nxv * t.wa) * 4u;
}
__device__ __forceinline__ CsView tile_view(const int *dst, const TileWin &t) { // virtual origin: view.p[(x) * rs + y] for absolute x, y
    CsView v;
    v.p = dst - (long long)t.lx * t.wa - t.y0a;
    v.rs = t.wa;
    return v;
}

extern __shared__ __align__(16) unsigned char ogn_dyn_smem[];

// ---------------------------------------------------------------------------------------------------------------
// Stand-alone kernels: one launch per layer phase.
template <int Z, int NCH, int W>
__global__ void __launch_bounds__(kThreads)
k_sc_forward_v(const __grid_constant__ ogn_fwd_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int *__restrict__ out0,
               int *__restrict__ out1) {
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) / W;
    if (pack >= npacks) return; // whole octets (W = 8) / warps (W = 32) leave together
    sc_forward_pack<Z, NCH, W, CsNc>(A, Hx, Hy, x0, pack, packsPerRow, blockIdx.y, out0, out1, nullptr);
}

template <int Z, int NCH, int MINB, int W>
__global__ void __launch_bounds__(kThreads, MINB)
k_pred_step_v(const __grid_constant__ ogn_pred_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int learn, int activate,
              const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    const int member = blockIdx.y;
    if (activate == 1) pred_copy_inputs<CsNc>(A, member, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) / W;
    if (pack >= npacks) return;
    pred_step_pack<Z, NCH, W, CsNc>(A, Hx, Hy, x0, pack, packsPerRow, member, learn, activate, target, alpha, acts, hiddenCs, nullptr);
}

// Predictor step with the weight rows staged in shared memory (8 * NS rows per lane and round; dynamic shared memory =
// blockDim.x * 8 * NS * 16 bytes)
template <int Z, int NS>
__global__ void __launch_bounds__(kThreads)
k_pred_step_a(const __grid_constant__ ogn_pred_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int learn, int activate,
              const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    const int member = blockIdx.y;
    if (activate == 1) pred_copy_inputs<CsNc>(A, member, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) / 8;
    if (pack >= npacks) return;
    float4 *stage = reinterpret_cast<float4 *>(ogn_dyn_smem) + threadIdx.x;
    pred_step_pack<Z, 1, 8, CsNc, NS>(A, Hx, Hy, x0, pack, packsPerRow, member, learn, activate, target, alpha, acts, hiddenCs, nullptr, stage);
}

template <int Z, int NCH, int W>
__global__ void __launch_bounds__(kThreads)
k_sc_recon_v(const __grid_constant__ ogn_learn_args A, const int *__restrict__ hcsAll, int vx0, int npacks, int packsPerRow,
             int2 *__restrict__ work, unsigned *workCount) {
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) / W;
    if (pack >= npacks) return;
    const int member = blockIdx.y;
    sc_recon_pack<Z, NCH, W, CsNc>(A, blockIdx.z, hcsAll, vx0, pack, packsPerRow, member, member * (int)gridDim.z + (int)blockIdx.z, work,
                                   workCount);
}

// TMA-staged variants of the two row-gather kernels (W = 8). ntiles CSDR windows of at most tileInts ints each are staged
// at the head of dynamic shared memory, the mbarrier sits behind them. All packs of a block lie in one hidden row
// (packsPerRow % (blockDim/8) == 0, checked by the launcher).
template <int Z, int NCH>
__global__ void __launch_bounds__(kThreads)
k_sc_forward_t(const __grid_constant__ ogn_fwd_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int tileInts,
               int *__restrict__ out0, int *__restrict__ out1) {
    int *tiles = reinterpret_cast<int *>(ogn_dyn_smem);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(tiles + (size_t)A.nvl * tileInts);
    __shared__ CsView views[OGN_MAX_VL];
    const int member = blockIdx.y;
    const int ppb = blockDim.x >> 3, pack0 = blockIdx.x * ppb;
    const int ox = x0 + pack0 / packsPerRow, q0 = pack0 % packsPerRow, q1 = min(q0 + ppb, packsPerRow) - 1;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        unsigned bytes = 0;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo];
            const TileWin t = tile_window(rf, ox, q0, q1);
            views[v] = tile_view(tiles + (size_t)v * tileInts, t);
            bytes += (unsigned)(t.nxv * t.wa) * 4u;
        }
        mbar_expect_tx(bar, bytes);
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo];
            const TileWin t = tile_window(rf, ox, q0, q1);
            tile_issue(tiles + (size_t)v * tileInts, A.vl[v].cs + (long long)member * rf.Vx * rf.Vy, rf.Vy, t, bar);
        }
    }
    __syncthreads(); // barrier initialised, views published
    mbar_wait(bar, 0);
    const int pack = pack0 + (threadIdx.x >> 3);
    if (pack >= npacks) return;
    sc_forward_pack<Z, NCH, 8, CsSm>(A, Hx, Hy, x0, pack, packsPerRow, member, out0, out1, views);
}

template <int Z, int NCH, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_pred_step_t(const __grid_constant__ ogn_pred_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int tileInts, int learn,
              int activate, const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    int *tiles = reinterpret_cast<int *>(ogn_dyn_smem);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(tiles + (size_t)2 * A.nvl * tileInts);
    __shared__ CsView views[2 * OGN_MAX_PRED_VL];
    const int member = blockIdx.y;
    if (activate == 1) pred_copy_inputs<CsNc>(A, member, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    const int ppb = blockDim.x >> 3, pack0 = blockIdx.x * ppb;
    const int ox = x0 + pack0 / packsPerRow, q0 = pack0 % packsPerRow, q1 = min(q0 + ppb, packsPerRow) - 1;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        unsigned bytes = 0;
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo];
            const TileWin t = tile_window(rf, ox, q0, q1);
            const unsigned b = (unsigned)(t.nxv * t.wa) * 4u;
            if (activate) { views[v] = tile_view(tiles + (size_t)(2 * v) * tileInts, t); bytes += b; }
            if (learn) { views[OGN_MAX_PRED_VL + v] = tile_view(tiles + (size_t)(2 * v + 1) * tileInts, t); bytes += b; }
        }
        mbar_expect_tx(bar, bytes);
        for (int v = 0; v < A.nvl; ++v) {
            const ogn_rf &rf = A.geo[A.vl[v].geo];
            const TileWin t = tile_window(rf, ox, q0, q1);
            const long long mo = (long long)member * rf.Vx * rf.Vy;
            if (activate) tile_issue(tiles + (size_t)(2 * v) * tileInts, A.vl[v].cs + mo, rf.Vy, t, bar);
            if (learn) tile_issue(tiles + (size_t)(2 * v + 1) * tileInts, A.vl[v].cs_prev + mo, rf.Vy, t, bar);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int pack = pack0 + (threadIdx.x >> 3);
    if (pack >= npacks) return;
    // the target CSDR is not staged (one read per column): CsSm::ld is a plain load, valid for global memory too
    pred_step_pack<Z, NCH, 8, CsSm>(A, Hx, Hy, x0, pack, packsPerRow, member, learn, activate, target, alpha, acts, hiddenCs, views);
}

// K3 part 2 as a persistent grid over the work units; the last warp to finish resets the list for the next launch.
template <int Z, int VER>
__global__ void __launch_bounds__(kThreads)
k_sc_update_v(const __grid_constant__ ogn_learn_args A, const int *__restrict__ hcsAll, const int *__restrict__ hcsPrevAll,
              const int2 *__restrict__ work, unsigned *workCount, unsigned *doneCount, float alpha, unsigned long long *updCount,
              int chunksPerEntry, int batches) {
    const int ITEMS = VER == 3 ? UpdV2<Z, 16>::ITEMS * batches : VER == 2 ? UpdV2<Z, 8>::ITEMS * batches : UpdV1<Z>::ITEMS;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
    const unsigned n = *(volatile unsigned *)workCount;
    const unsigned long long nunits = (unsigned long long)n * (unsigned)chunksPerEntry;
    int nupd = 0;
    for (unsigned long long unit = warp; unit < nunits; unit += nwarps) {
        const unsigned e = (unsigned)(unit / (unsigned)chunksPerEntry);
        const int k0 = (int)(unit - (unsigned long long)e * (unsigned)chunksPerEntry) * ITEMS;
        const int2 wk = work[e];
        if constexpr (VER == 3) nupd += sc_update_unit_v2<Z, 16, CsNc, false>(A, hcsAll, hcsPrevAll, wk, k0, batches, alpha);
        else if constexpr (VER == 2) nupd += sc_update_unit_v2<Z, 8, CsNc, false>(A, hcsAll, hcsPrevAll, wk, k0, batches, alpha);
        else nupd += sc_update_unit_v1<Z, CsNc, false>(A, hcsAll, hcsPrevAll, wk, k0, alpha);
    }
    // bookkeeping: pairs rewritten, and the last BLOCK out empties the list (one atomic per block: a launch with little work
    // used to spend most of its time on 9472 warps queueing at one L2 atomic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nupd += __shfl_xor_sync(0xffffffffu, nupd, o);
    if (lane == 0 && updCount && nupd) atomicAdd(updCount, (unsigned long long)nupd);
    __syncthreads(); // every warp of the block has read the list for the last time
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(doneCount, 1u) == gridDim.x - 1u) { *workCount = 0; *doneCount = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The whole step of a SMALL hierarchy in ONE launch (ogn_hier_step_small): the reference pays ~54
// fork/joins per step (Helpers.cpp:15-72, Hierarchy.cpp:219-284) and the per-layer kernels above 8-16 launches; here the
// layer phases of a step (SparseCoder forward, reconstruction, update, Predictor learn+activate, in the order
// Hierarchy::step prescribes) are the iterations of a loop inside one kernel, separated by a barrier among the blocks
// that work on the same hierarchy. grid = (blocks per member, ensemble members): members never wait for each other, so an
// ensemble (BASELINE config 5) streams its weights with every member at its own phase, and a single small hierarchy
// (configs 1-3) runs on a cooperative grid with one software barrier per phase. CSDRs produced by one phase are read by
// the next through L2 (CsCg).
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// barrier among the nb blocks of one member; bar = {arrivals, generation} of that member
__device__ __forceinline__ void member_barrier(unsigned *bar, int nb) {
    __syncthreads();
    if (nb > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned gen = ld_acquire_gpu(bar + 1);
            if (atomicAdd(bar, 1u) == (unsigned)nb - 1u) {
                *bar = 0;
                __threadfence();
                st_release_gpu(bar + 1, gen + 1u);
            } else {
                int spins = 0;
                while (ld_acquire_gpu(bar + 1) == gen)
                    if (++spins > (1 << 23)) __trap(); // a lost block is a bug: fail the launch instead of hanging
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// One phase body per (kind, row width, mapping), compiled as separate functions (noinline) so that each keeps the register
// allocation of its stand-alone kernel instead of the union of all of them.
template <int Z, int W, class CS>
__device__ __noinline__ void small_fwd(const ogn_small_phase &P, int member, int unit0, int stride) {
    for (int pack = unit0; pack < P.npacks; pack += stride)
        sc_forward_pack<Z, 1, W, CS>(P.u.fwd, P.Hx, P.Hy, 0, pack, P.ppr, member, P.out0, P.out1, nullptr);
}
template <int Z, int W, class CS>
__device__ __noinline__ void small_recon(const ogn_small_phase &P, int member, int unit0, int stride, unsigned *cnt) {
    int2 *work = reinterpret_cast<int2 *>(P.work) + (long long)member * P.workStride;
    const int total = P.npacks * P.u.learn.nvl; // (visible layer, pack) pairs
    for (int it = unit0; it < total; it += stride) {
        const int vli = it / P.npacks, pack = it - vli * P.npacks;
        // the tag keeps the stand-alone encoding member * nvl + vli so that the update units decode it the same way
        sc_recon_pack<Z, 1, W, CS>(P.u.learn, vli, P.hcs, 0, pack, P.ppr, member, member * P.u.learn.nvl + vli, work, cnt);
    }
}
template <int Z, int VER, class CS>
__device__ __noinline__ void small_update(const ogn_small_phase &P, int member, int warp, int nwarps, unsigned *cnt) {
    const int ITEMS = VER == 2 ? UpdV2<Z>::ITEMS * P.updBatches : UpdV1<Z>::ITEMS;
    const int2 *work = reinterpret_cast<const int2 *>(P.work) + (long long)member * P.workStride;
    const unsigned n = __ldcg(cnt);
    const long long nunits = (long long)n * P.chunks;
    int nupd = 0;
    for (long long unit = warp; unit < nunits; unit += nwarps) {
        const unsigned e = (unsigned)(unit / P.chunks);
        const int c = (int)(unit - (long long)e * P.chunks);
        const int2 wk = __ldcg(work + e);
        if constexpr (VER == 2) nupd += sc_update_unit_v2<Z, 8, CS, true>(P.u.learn, P.hcs, P.hcsPrev, wk, c * ITEMS, P.updBatches, P.alpha);
        else nupd += sc_update_unit_v1<Z, CS, true>(P.u.learn, P.hcs, P.hcsPrev, wk, c * ITEMS, P.alpha);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nupd += __shfl_xor_sync(0xffffffffu, nupd, o);
    if ((threadIdx.x & 31) == 0 && P.updCount && nupd) atomicAdd(P.updCount, (unsigned long long)nupd);
}
template <int Z, int W, class CS>
__device__ __noinline__ void small_pred(const ogn_small_phase &P, int member, int unit0, int stride) {
    for (int pack = unit0; pack < P.npacks; pack += stride)
        pred_step_pack<Z, 1, W, CS>(P.u.pred, P.Hx, P.Hy, 0, pack, P.ppr, member, P.learn, P.activate, P.target, P.alpha, P.acts, P.out0,
                                      nullptr);
}

#define OGN_SMALL_Z(CALL8, CALL16, CALL32) \
    switch (P.z) {                         \
    case 8: CALL8; break;                  \
    case 16: CALL16; break;                \
    default: CALL32; break;                \
    }

// counters: per member OGN_SMALL_COUNTERS words: [0..1] barrier {arrivals, generation}, [2 + i] fill of work list i. All are
// zero between launches. W = lanes per pack of every phase of the launch (8: ensembles, 32: a single small hierarchy). CS =
// how CSDRs produced earlier in the launch are read: plain cached loads when ONE block owns the member (an SM's L1 is
// coherent with its own stores), L2 loads (CsCg) when several blocks share it. phaseBytes > 0: the phase array is first
// copied into dynamic shared memory, so that every descriptor read of the phase loop is an LDS instead of an L2 access.
template <int W, class CS>
__global__ void __launch_bounds__(512, 1)
k_hier_step_small(const ogn_small_phase *__restrict__ phases, int nphases, const __grid_constant__ ogn_small_inputs I, unsigned *counters,
                  int countersPerMember, int phaseBytes, unsigned long long *trace) {
    const int member = blockIdx.y, nb = gridDim.x;
    // debug timeline (ogn_debug_small_trace): thread 0 of blocks 0 and nb - 1 of member 0 stamps globaltimer at the start of
    // every phase, after it and after its barrier: trace[(block ? 1 : 0) * 3 * 256 + 3 * ph + {0, 1, 2}]
    const bool tr = trace && threadIdx.x == 0 && blockIdx.y == 0 && (blockIdx.x == 0 || (int)blockIdx.x == nb - 1);
    unsigned long long *tb = trace + (blockIdx.x ? 768 : 0);
#define SMALL_STAMP(i) do { if (tr && ph < 256) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tb[3 * ph + (i)] = t_; } } while (0)
    unsigned *mc = counters + (long long)member * countersPerMember;
    if (phaseBytes > 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(phases);
        uint4 *dst = reinterpret_cast<uint4 *>(ogn_dyn_smem);
        for (int i = threadIdx.x; i < phaseBytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
        phases = reinterpret_cast<const ogn_small_phase *>(ogn_dyn_smem);
    }
    if (I.n > 0) { // K1: every block copies its member's inputs itself (identical values), so no block waits for another
        for (int i = 0; i < I.n; ++i) {
            const long long mo = (long long)member * I.count[i];
            for (int k = threadIdx.x; k < I.count[i]; k += blockDim.x) I.dst[i][mo + k] = __ldg(I.src[i] + mo + k);
        }
    }
    __syncthreads();
    // units of this block: octets (W = 8) or warps (W = 32), strided over the member's blocks
    const int unitsPerBlock = blockDim.x / W, unit0 = blockIdx.x * unitsPerBlock + threadIdx.x / W, stride = nb * unitsPerBlock;
    const int warpsPerBlock = blockDim.x >> 5, nwarps = nb * warpsPerBlock;
    for (int ph = 0; ph < nphases; ++ph) {
        const ogn_small_phase &P = phases[ph];
        unsigned *cnt = mc + 2 + P.counter;
        SMALL_STAMP(0);
        // the phases between two barriers are independent: each starts on its own range of the launch's unit slots
        // (unitShift), so a slot that has nothing to do in one phase is already working on the next one
        int u0 = unit0 - P.unitShift % stride;
        if (u0 < 0) u0 += stride;
        int w0 = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5) - P.unitShift % nwarps;
        if (w0 < 0) w0 += nwarps;
        switch (P.kind) {
        case OGN_PHASE_SC_FORWARD:
            OGN_SMALL_Z((small_fwd<8, W, CS>(P, member, u0, stride)), (small_fwd<16, W, CS>(P, member, u0, stride)),
                        (small_fwd<32, W, CS>(P, member, u0, stride)));
            break;
        case OGN_PHASE_SC_RECON:
            OGN_SMALL_Z((small_recon<8, W, CS>(P, member, u0, stride, cnt)), (small_recon<16, W, CS>(P, member, u0, stride, cnt)),
                        (small_recon<32, W, CS>(P, member, u0, stride, cnt)));
            break;
        case OGN_PHASE_SC_UPDATE:
            if (P.updVer == 2) {
                OGN_SMALL_Z((small_update<8, 2, CS>(P, member, w0, nwarps, cnt)), (small_update<16, 2, CS>(P, member, w0, nwarps, cnt)),
                            (small_update<32, 2, CS>(P, member, w0, nwarps, cnt)));
            } else {
                OGN_SMALL_Z((small_update<8, 1, CS>(P, member, w0, nwarps, cnt)), (small_update<16, 1, CS>(P, member, w0, nwarps, cnt)),
                            (small_update<32, 1, CS>(P, member, w0, nwarps, cnt)));
            }
            break;
        case OGN_PHASE_PRED:
            if (P.activate == 1) pred_copy_inputs<CS>(P.u.pred, member, blockIdx.x * blockDim.x + threadIdx.x, nb * blockDim.x);
            OGN_SMALL_Z((small_pred<8, W, CS>(P, member, u0, stride)), (small_pred<16, W, CS>(P, member, u0, stride)),
                        (small_pred<32, W, CS>(P, member, u0, stride)));
            break;
        default: break;
        }
        SMALL_STAMP(1);
        if (ph + 1 < nphases && P.barrierAfter) member_barrier(mc, nb);
        SMALL_STAMP(2);
    }
#undef SMALL_STAMP
    // the work-list fills go back to zero for the next launch: a reconstruction is always followed by its update behind a
    // barrier, so after one more barrier nobody reads a fill any more
    member_barrier(mc, nb);
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(countersPerMember - 2)) mc[2 + threadIdx.x] = 0;
}

} // namespace vec
