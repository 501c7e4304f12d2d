// ogn_vec.cuh — the 128-bit ("octet") kernels of the Sparse Predictive Hierarchy step: the fast path taken whenever a
// gathered row holds 8, 16 or 32 cells, i.e. whenever the packed F / R layouts of ogn_kernels_api.h put exactly one
// 128-byte line behind every (column pack, receptive-field slot, active input cell). Included by ogn_kernels.cu.
//
// Execution model: an OCTET of 8 lanes owns one 128-byte line; lane `sub` holds floats [4*sub, 4*sub+4) of it, which
// belong to column p = 4*sub / Z of the pack and are that column's cells (4*sub % Z) .. +3. A warp is four independent
// octets = four packs, so one LDG.128 / STG.128 instruction moves four full 128-byte lines (512 B) and a batch of
// 8*NCH of them is issued before the first dependent add: 4*NCH KB in flight per warp, 4x the bytes and a quarter of the
// instructions of the scalar kernels. The lanes of an octet take turns turning "their" slot of the next chunk of 8 slots
// into a line descriptor (one CSDR read, prefetched one chunk ahead so that its L2 latency hides behind the weight loads
// in flight); descriptors travel by width-8 shuffles. Every cell is still accumulated by ONE lane in the reference's
// slot order (loads are batched, adds are sequential), so sums stay bit-identical to SparseMatrix::multiplyOHVs /
// multiplyOHVsT (SparseMatrix.cpp:260-294).
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
__device__ __forceinline__ void st_plain(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
template <bool kStream> __device__ __forceinline__ float4 ld_w(const float *p) { return kStream ? ld_stream(p) : ld_plain(p); }

__device__ __forceinline__ unsigned octet_mask() { return 0xffu << (threadIdx.x & 24); }

// Walks the union field of one pack in chunks of 8 slots; `next()` returns this lane's descriptor of its slot of the
// next chunk: (slot * Vz + active cell) | dyU << 20 (dy bits = kNoDy for the clamped slots past the end of the field).
struct SlotWalker {
    const int *cs;       // visible CSDR of this member
    int lx, lyU, nyU, nxv, n, Vy, Vz;
    int s, dx, dy, cv;   // this lane's slot of the chunk being described and its prefetched CSDR value
    __device__ __forceinline__ void fetch() {
        const bool in = s < n;
        cv = __ldg(cs + (lx + (in ? dx : nxv - 1)) * Vy + lyU + (in ? dy : nyU - 1));
    }
    __device__ __forceinline__ void init(const int *cs_, const PackCtx &c, int Vy_, int Vz_, int sub) {
        cs = cs_; lx = c.lx; lyU = c.lyU; nyU = c.nyU; n = c.nslotsU; nxv = n / max(nyU, 1); Vy = Vy_; Vz = Vz_;
        s = sub; dx = s / max(nyU, 1); dy = s - dx * nyU;
        fetch();
    }
    __device__ __forceinline__ unsigned next() {
        const bool in = s < n;
        const unsigned d = (unsigned)((in ? s : n - 1) * Vz + cv) | ((in ? (unsigned)dy : kNoDy) << kLineBits);
        s += 8; dy += 8;
        while (dy >= nyU) { dy -= nyU; ++dx; }
        fetch();
        return d;
    }
};

// Σ over the lane's column's field of its 4 cells of W_F[...][cs[visible column]] (multiplyOHVs, SparseMatrix.cpp:260-276).
// wl = pack block + 4*sub. NCH chunks of 8 line loads are issued per batch.
template <bool kStream, int NCH>
__device__ __forceinline__ float4 gather(unsigned om, int sub, const float *wl, const int *cs, const PackCtx &c, int Vy, int Vz) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c.nslotsU <= 0) return acc;
    SlotWalker sw;
    sw.init(cs, c, Vy, Vz, sub);
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
    return acc;
}

// W_F[...][cs[visible column]] += delta over the lane's column's field (deltaOHVs, SparseMatrix.cpp:488-501).
template <int NCH>
__device__ __forceinline__ void add_delta(unsigned om, int sub, float *wl, const int *cs, const float4 &delta, const PackCtx &c, int Vy,
                                          int Vz) {
    if (c.nslotsU <= 0) return;
    SlotWalker sw;
    sw.init(cs, c, Vy, Vz, sub);
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

// ---------------------------------------------------------------------------------------------------------------
// cp.async ring (template parameter D > 0 of the kernels): instead of holding a batch of lines in registers, every lane
// copies "its" 16 bytes of each line straight into a private slice of shared memory (cp.async.cg, L2 -> smem, no
// registers) and keeps D chunks of 8 lines in flight; a chunk is summed, in slot order, as soon as it has landed and its
// stage is refilled with the chunk D ahead. One warp therefore has up to D * 4 KB in flight on its own, so a column's
// whole field (11 chunks at radius 4) is covered by ~2 memory round trips instead of 11: short waves, no tail when a
// sharded launch has only a couple of waves of work. Lanes only ever read back bytes they copied themselves, so
// cp.async.wait_group is the only synchronisation. Ring layout per warp: [stage][slot j][lane] float4 (conflict-free).
__device__ __forceinline__ void cp_async16(float4 *smem, const float *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Like SlotWalker, but the CSDR read of a chunk is handed out as a (meta, value) pair whose value the caller consumes
// later: the reads of the first D chunks go out together instead of one L2 round trip each. Descriptor = meta + cv.
struct SlotWalkerP {
    struct Pending { unsigned meta; int cv; };
    const int *cs;
    int lx, lyU, nyU, nxv, n, Vy, Vz;
    int s, dx, dy;
    __device__ __forceinline__ void init(const int *cs_, const PackCtx &c, int Vy_, int Vz_, int sub) {
        cs = cs_; lx = c.lx; lyU = c.lyU; nyU = c.nyU; n = c.nslotsU; nxv = n / max(nyU, 1); Vy = Vy_; Vz = Vz_;
        s = sub; dx = s / max(nyU, 1); dy = s - dx * nyU;
    }
    __device__ __forceinline__ Pending fetch() {
        const bool in = s < n;
        Pending p;
        p.meta = (unsigned)((in ? s : n - 1) * Vz) | ((in ? (unsigned)dy : kNoDy) << kLineBits);
        p.cv = __ldg(cs + (lx + (in ? dx : nxv - 1)) * Vy + lyU + (in ? dy : nyU - 1));
        s += 8; dy += 8;
        while (dy >= nyU) { dy -= nyU; ++dx; }
        return p;
    }
};

// queue the 8 line copies of one chunk into `stage` (already offset by the lane); returns the own-field mask of the chunk
__device__ __forceinline__ unsigned issue_chunk(unsigned om, float4 *stage, const float *wl, unsigned d, const PackCtx &c) {
    unsigned vm = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const unsigned dj = __shfl_sync(om, d, j, 8);
        vm |= ((unsigned)((int)(dj >> kLineBits) - c.dlo) < (unsigned)c.dn ? 1u : 0u) << j;
        cp_async16(stage + j * 32, wl + ((size_t)(dj & kLineMask) << 5));
    }
    return vm;
}

template <int D>
__device__ __forceinline__ float4 gather_p(unsigned om, int sub, float4 *ring, const float *wl, const int *cs, const PackCtx &c, int Vy,
                                           int Vz) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (c.nslotsU <= 0) return acc;
    SlotWalkerP sw;
    sw.init(cs, c, Vy, Vz, sub);
    const int nch = (c.nslotsU + 7) >> 3;
    unsigned vm[D];
    SlotWalkerP::Pending pq[D];
#pragma unroll
    for (int st = 0; st < D; ++st) pq[st] = sw.fetch();
    SlotWalkerP::Pending nxt = sw.fetch();
#pragma unroll
    for (int st = 0; st < D; ++st) {
        vm[st] = 0;
        if (st < nch) vm[st] = issue_chunk(om, ring + st * 256, wl, pq[st].meta + (unsigned)pq[st].cv, c);
        cp_async_commit();
    }
    for (int ch0 = 0; ch0 < nch; ch0 += D) {
#pragma unroll
        for (int st = 0; st < D; ++st) {
            const int ch = ch0 + st;
            if (ch < nch) {
                cp_async_wait<D - 1>();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = ring[(st * 8 + j) * 32];
                    if ((vm[st] >> j) & 1u) { acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w; }
                }
                if (ch + D < nch) {
                    vm[st] = issue_chunk(om, ring + st * 256, wl, nxt.meta + (unsigned)nxt.cv, c);
                    nxt = sw.fetch();
                }
                cp_async_commit();
            }
        }
    }
    return acc;
}

template <int D>
__device__ __forceinline__ void add_delta_p(unsigned om, int sub, float4 *ring, float *wl, const int *cs, const float4 &delta,
                                            const PackCtx &c, int Vy, int Vz) {
    if (c.nslotsU <= 0) return;
    SlotWalkerP sw;
    sw.init(cs, c, Vy, Vz, sub);
    const int nch = (c.nslotsU + 7) >> 3;
    unsigned dq[D];
    SlotWalkerP::Pending pq[D];
#pragma unroll
    for (int st = 0; st < D; ++st) pq[st] = sw.fetch();
    SlotWalkerP::Pending nxt = sw.fetch();
#pragma unroll
    for (int st = 0; st < D; ++st) {
        dq[st] = pq[st].meta + (unsigned)pq[st].cv;
        if (st < nch) (void)issue_chunk(om, ring + st * 256, wl, dq[st], c);
        cp_async_commit();
    }
    for (int ch0 = 0; ch0 < nch; ch0 += D) {
#pragma unroll
        for (int st = 0; st < D; ++st) {
            const int ch = ch0 + st;
            if (ch < nch) {
                cp_async_wait<D - 1>();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned dj = __shfl_sync(om, dq[st], j, 8);
                    if ((unsigned)((int)(dj >> kLineBits) - c.dlo) < (unsigned)c.dn) {
                        float4 v = ring[(st * 8 + j) * 32];
                        v.x += delta.x; v.y += delta.y; v.z += delta.z; v.w += delta.w;
                        st_plain(wl + ((size_t)(dj & kLineMask) << 5), v);
                    }
                }
                if (ch + D < nch) {
                    dq[st] = nxt.meta + (unsigned)nxt.cv;
                    (void)issue_chunk(om, ring + st * 256, wl, dq[st], c);
                    nxt = sw.fetch();
                }
                cp_async_commit();
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

// ---------------------------------------------------------------------------------------------------------------
// K2: SparseCoder::forward (SparseCoder.cpp:13-43), Hz = Z in {8, 16, 32}.
extern __shared__ float4 ogn_ring[]; // cp.async rings of the block's warps (D > 0): [warp][stage][slot][lane]

template <int Z, int NCH, int D>
__global__ void __launch_bounds__(kThreads)
k_sc_forward_v(const __grid_constant__ ogn_fwd_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int *__restrict__ out0,
               int *__restrict__ out1) {
    const int sub = threadIdx.x & 7;
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (pack >= npacks) return; // whole octets leave together
    const unsigned om = octet_mask();
    const int member = blockIdx.y;
    const PackId id = pack_id<Z>(pack, packsPerRow, x0, Hy, sub);
    float4 *ring = ogn_ring + (threadIdx.x >> 5) * (D * 256) + (threadIdx.x & 31);

    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_fwd_vl &vl = A.vl[v];
        const ogn_rf &rf = A.geo[vl.geo];
        const PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
        const float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
        const int *cs = vl.cs + (long long)member * rf.Vx * rf.Vy;
        float4 g;
        if constexpr (D > 0) g = gather_p<D>(om, sub, ring, wl, cs, c, rf.Vy, rf.Vz);
        else g = gather<true, NCH>(om, sub, wl, cs, c, rf.Vy, rf.Vz);
        sum.x += g.x; sum.y += g.y; sum.z += g.z; sum.w += g.w;
    }
    const int best = column_argmax_fwd<Z>(om, id.hz0, sum);
    if (id.hz0 == 0 && id.colValid) {
        const long long o = (long long)member * Hx * Hy + id.ox * Hy + id.oy;
        out0[o] = best;
        if (out1) out1[o] = best;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K6+K7+K8: Predictor::learn (Predictor.cpp:49-70), Predictor::forward (:13-47), inputCs -> inputCsPrev (:118-126),
// Hz = Z in {8, 16, 32}.
template <int Z, int NCH, int MINB, int D>
__global__ void __launch_bounds__(kThreads, MINB)
k_pred_step_v(const __grid_constant__ ogn_pred_args A, int Hx, int Hy, int x0, int npacks, int packsPerRow, int learn, int activate,
              const int *__restrict__ target, float alpha, float *acts, int *__restrict__ hiddenCs) {
    const int member = blockIdx.y;
    if (activate == 1) { // K8: the grid copies the inputs into the *other* inputCsPrev buffer (never read by this launch)
        const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
        for (int v = 0; v < A.nvl; ++v) {
            if (!A.vl[v].cs_prev_out) continue;
            const int n = A.geo[A.vl[v].geo].Vx * A.geo[A.vl[v].geo].Vy;
            const long long mo = (long long)member * n;
            for (int i = tid; i < n; i += nth) A.vl[v].cs_prev_out[mo + i] = __ldg(A.vl[v].cs + mo + i);
        }
    }
    const int sub = threadIdx.x & 7;
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (pack >= npacks) return;
    const unsigned om = octet_mask();
    const PackId id = pack_id<Z>(pack, packsPerRow, x0, Hy, sub);
    const long long col = (long long)member * Hx * Hy + id.ox * Hy + (id.colValid ? id.oy : Hy - 1);
    float *act4 = acts + col * Z + id.hz0;
    float4 *ring = ogn_ring + (threadIdx.x >> 5) * (D * 256) + (threadIdx.x & 31);

    if (learn) {
        const int t = __ldg(target + col);
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
            const int *cs = vl.cs_prev + (long long)member * rf.Vx * rf.Vy;
            if constexpr (D > 0) add_delta_p<D>(om, sub, ring, wl, cs, delta, c, rf.Vy, rf.Vz);
            else add_delta<NCH>(om, sub, wl, cs, delta, c, rf.Vy, rf.Vz);
        }
        if (D > 0) __threadfence(); // the activate pass re-reads lines this lane just stored, through cp.async
    }
    if (!activate) return;

    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int count = 0;
    for (int v = 0; v < A.nvl; ++v) {
        const ogn_pred_vl &vl = A.vl[v];
        const ogn_rf &rf = A.geo[vl.geo];
        const PackCtx c = pack_ctx(rf, id.ox, id.q, id.p);
        const float *wl = vl.wf + (c.blk + (long long)member * vl.mstride) + sub * 4;
        const int *cs = vl.cs + (long long)member * rf.Vx * rf.Vy;
        float4 g;
        if constexpr (D > 0) g = gather_p<D>(om, sub, ring, wl, cs, c, rf.Vy, rf.Vz);
        else g = gather<false, NCH>(om, sub, wl, cs, c, rf.Vy, rf.Vz);
        sum.x += g.x; sum.y += g.y; sum.z += g.z; sum.w += g.w;
        count += c.dn * (c.nslotsU / max(c.nyU, 1)); // own field: nx * ny
    }
    const float cnt = (float)count;
    sum.x = div_ref(sum.x, cnt); sum.y = div_ref(sum.y, cnt); sum.z = div_ref(sum.z, cnt); sum.w = div_ref(sum.w, cnt);
    if (id.colValid) st_plain(act4, sum);
    const int best = column_argmax_fwd<Z>(om, id.hz0, sum);
    if (id.hz0 == 0 && id.colValid) hiddenCs[col] = best;
}

// ---------------------------------------------------------------------------------------------------------------
// K3, part 1: the reconstruction of SparseCoder::learn (SparseCoder.cpp:45-70, multiplyOHVsT SparseMatrix.cpp:278-294,
// countT :215-221), Vz = Z in {8, 16, 32}. One octet per pack of 32/Z y-adjacent visible columns; lane = 4 visible
// cells. Visible columns whose reconstruction arg-max misses the target are appended to a work list for part 2.
struct CoverWalker {
    const int *hcs;
    int hx0, hy0U, ncyU, n, vx, vyBase;
    int k, ix, iy; // this lane's item of the chunk being described: hidden column (hx0 + ix, hy0U + iy)
    template <int Z> __device__ __forceinline__ unsigned long long describe(const ogn_rf &rf) const {
        constexpr int PR = 32 / Z;
        const bool in = k < n;
        const int ox = hx0 + (in ? ix : (n - 1) / ncyU), oy = hy0U + (in ? iy : (n - 1) % ncyU);
        const int hc = __ldg(hcs + ox * rf.Hy + oy);
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
        k += 8; iy += 8;
        while (iy >= ncyU) { iy -= ncyU; ++ix; }
    }
};

template <int Z, int NCH, int D>
__global__ void __launch_bounds__(kThreads)
k_sc_recon_v(const __grid_constant__ ogn_learn_args A, const int *__restrict__ hcsAll, int vx0, int npacks, int packsPerRow,
             int2 *__restrict__ work, unsigned *workCount) {
    constexpr int PR = 32 / Z;
    const int sub = threadIdx.x & 7;
    const int pack = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (pack >= npacks) return;
    const unsigned om = octet_mask();
    const int member = blockIdx.y;
    const ogn_learn_vl &vl = A.vl[blockIdx.z];
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
        CoverWalker cw;
        cw.hcs = hcsAll + (long long)member * rf.Hx * rf.Hy;
        cw.hx0 = hx0; cw.hy0U = hy0U; cw.ncyU = ncyU; cw.n = ncovU; cw.vx = vx; cw.vyBase = vyBase;
        cw.k = sub; cw.ix = sub / ncyU; cw.iy = sub - cw.ix * ncyU;
        if constexpr (D > 0) {
            // cp.async ring: D chunks of 8 lines in flight per lane group (see gather_p)
            float4 *ring = ogn_ring + (threadIdx.x >> 5) * (D * 256) + (threadIdx.x & 31);
            const int nch = (ncovU + 7) >> 3;
            unsigned long long dq[D];
            unsigned vm[D];
#pragma unroll
            for (int st = 0; st < D; ++st) { dq[st] = cw.describe<Z>(rf); cw.advance(); }
            unsigned long long nxt = cw.describe<Z>(rf);
            cw.advance();
            auto issue = [&](int st, unsigned long long d) {
                unsigned m = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned long long dj = __shfl_sync(om, d, j, 8);
                    m |= (unsigned)((dj >> pv) & 1ull) << j;
                    cp_async16(ring + (st * 8 + j) * 32, wr + (dj & ~31ull));
                }
                return m;
            };
#pragma unroll
            for (int st = 0; st < D; ++st) {
                vm[st] = 0;
                if (st < nch) vm[st] = issue(st, dq[st]);
                cp_async_commit();
            }
            for (int ch0 = 0; ch0 < nch; ch0 += D) {
#pragma unroll
                for (int st = 0; st < D; ++st) {
                    const int ch = ch0 + st;
                    if (ch < nch) {
                        cp_async_wait<D - 1>();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 w = ring[(st * 8 + j) * 32];
                            if ((vm[st] >> j) & 1u) { sum.x += w.x; sum.y += w.y; sum.z += w.z; sum.w += w.w; }
                        }
                        if (ch + D < nch) {
                            vm[st] = issue(st, nxt);
                            nxt = cw.describe<Z>(rf);
                            cw.advance();
                        }
                        cp_async_commit();
                    }
                }
            }
        } else {
        unsigned long long d[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) { d[k] = cw.describe<Z>(rf); cw.advance(); }
        for (int ch = 0; ch < ncovU; ch += 8 * NCH) {
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
            for (int k = 0; k < NCH; ++k) { d[k] = cw.describe<Z>(rf); cw.advance(); }
#pragma unroll
            for (int j = 0; j < 8 * NCH; ++j)
                if ((vm >> j) & 1u) { sum.x += w[j].x; sum.y += w[j].y; sum.z += w[j].z; sum.w += w[j].w; }
        }
        }
    }
    const float cnt = (float)ncov;
    sum.x = div_ref(sum.x, cnt); sum.y = div_ref(sum.y, cnt); sum.z = div_ref(sum.z, cnt); sum.w = div_ref(sum.w, cnt);
    if (colValid) st_plain(vl.recon + vcol * Z + vc0, sum);

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
    if (colValid && vc0 == 0 && i != __ldg(vl.cs + vcol)) {
        const unsigned slot = atomicAdd(workCount, 1u);
        work[slot] = make_int2(member * (int)gridDim.z + (int)blockIdx.z, vx * rf.Vy + vy);
    }
}

// K3, part 2: the delta rule on the weights of hidden columns whose state changed (deltaChangedOHVsT,
// SparseMatrix.cpp:572-590; SparseCoder.cpp:72-81) for the visible columns on the work list. A work unit is one chunk of
// 32/LC * UN cover items of one listed column, one warp per unit; LC = Z/4 lanes per (hidden column, visible column)
// pair, so 32/LC pairs are rewritten per warp instruction: the R copy with one 128-bit read-modify-write, the F copy with
// four scalar stores (one per forward row; 16 partially written sectors per pair — the HBM cost of keeping two layouts,
// measured in profiles/README.md: the kernel is DRAM-bound on those sector fills and write-backs, an L2 prefetch of the
// sectors made it slower). The last warp to finish resets the list for the next launch.

template <int Z>
__global__ void __launch_bounds__(kThreads)
k_sc_update_v(const __grid_constant__ ogn_learn_args A, const int *__restrict__ hcsAll, const int *__restrict__ hcsPrevAll,
              const int2 *__restrict__ work, unsigned *workCount, unsigned *doneCount, float alpha, unsigned long long *updCount,
              int chunksPerEntry) {
    constexpr int LC = Z / 4, IPW = 32 / LC, PR = 32 / Z, UN = 4;
    const ogn_rf &rf = A.geo;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
    const int it = lane / LC, vc0 = (lane % LC) * 4;
    const unsigned n = *(volatile unsigned *)workCount;
    const unsigned long long nunits = (unsigned long long)n * (unsigned)chunksPerEntry;
    const long long fstep = (long long)rf.pf * rf.Hz; // floats between the forward rows of consecutive visible cells
    int nupd = 0;
    for (unsigned long long unit = warp; unit < nunits; unit += nwarps) {
        const unsigned e = (unsigned)(unit / (unsigned)chunksPerEntry);
        const int k0 = (int)(unit - (unsigned long long)e * (unsigned)chunksPerEntry) * (IPW * UN);
        const int2 wk = work[e];
        const int vx = wk.y / rf.Vy, vy = wk.y - vx * rf.Vy;
        const int hx0 = __ldg(rf.rlox + vx), hx1 = __ldg(rf.rhix + vx);
        const int hy0 = __ldg(rf.rloy + vy), hy1 = __ldg(rf.rhiy + vy);
        const int ncy = max(0, hy1 - hy0 + 1), ncov = max(0, hx1 - hx0 + 1) * ncy;
        if (k0 >= ncov) continue;
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
                const int hc = __ldg(hcs + h);
                if (hc != __ldg(hcsPrev + h)) {
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
        const int target = __ldg(vl.cs + vcol);
        const float4 r = ld_plain(vl.recon + vcol * Z + vc0);
        float4 delta;
        delta.x = alpha * ((vc0 + 0 == target ? 1.0f : 0.0f) - ogn_expf(r.x));
        delta.y = alpha * ((vc0 + 1 == target ? 1.0f : 0.0f) - ogn_expf(r.y));
        delta.z = alpha * ((vc0 + 2 == target ? 1.0f : 0.0f) - ogn_expf(r.z));
        delta.w = alpha * ((vc0 + 3 == target ? 1.0f : 0.0f) - ogn_expf(r.w));
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
    }
    // bookkeeping: pairs rewritten, and the last warp out empties the list
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nupd += __shfl_xor_sync(0xffffffffu, nupd, o);
    if (lane == 0) {
        if (updCount && nupd) atomicAdd(updCount, (unsigned long long)nupd);
        __threadfence();
        if (atomicAdd(doneCount, 1u) == (unsigned)nwarps - 1u) { *workCount = 0; *doneCount = 0; __threadfence(); }
    }
}

} // namespace vec
