// tools/microbench_patterns.cu — what does a B200 deliver for the ACCESS PATTERNS of the two kernels that sit below the
// copy roofline, with the arithmetic, the CSDR reads and the receptive-field bookkeeping taken away?
//
//   rows:   the Predictor pattern (k_pred_step_v at config 4). A weight matrix of `packs` x `slots` x `vz` rows of 128 B;
//           per (pack, slot) ONE row is touched, chosen by a hash (the active cell of the input column). Pass "rmw" reads and
//           rewrites the chosen rows (learn), pass "gather" reads rows chosen by another hash (activate). Eight lanes own a
//           row (float4 each), `UN` rows in flight per lane, like the product kernel.
//   pairs:  the SparseCoder update pattern (k_sc_update_v). Per updated (visible column, hidden column) pair: a 64-byte row
//           of one matrix is read and rewritten, and 16 single floats 128 B apart (one per visible cell, inside one 2 KB tile)
//           are read and rewritten in the other matrix.
//
// Prints one JSON line per measurement: bytes the pattern touches (algorithmic), time per pass, GB/s. Standalone:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb tools/microbench_patterns.cu && /tmp/mb
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned mix(unsigned a, unsigned b, unsigned c) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ c * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

// cs != NULL: the chosen row comes from an int array in global memory (a dependent load, like the CSDR read of the product
// kernel) instead of from a hash. Dynamic shared memory only limits the resident blocks per SM.
template <int UN>
__global__ void __launch_bounds__(256) k_rows(float4 *w, long long packs, int slots, int vz, unsigned seedA, unsigned seedB, int mode, float *sink,
                                              const int *__restrict__ cs) {
    const int sub = threadIdx.x & 7;
    const long long pack = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
    if (pack >= packs) return;
    float4 *base = w + pack * slots * vz * 8 + sub;
    if (mode & 1) {
        for (int s = 0; s < slots; s += UN) {
            float4 v[UN];
            long long at[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const int ss = min(s + k, slots - 1);
                const unsigned pick = cs ? (unsigned)__ldg(cs + pack * slots + ss) : mix((unsigned)pack, (unsigned)ss, seedA) % (unsigned)vz;
                at[k] = ((long long)ss * vz + pick) * 8;
                v[k] = base[at[k]];
            }
#pragma unroll
            for (int k = 0; k < UN; ++k)
                if (s + k < slots) { v[k].x += 1e-9f; v[k].y += 1e-9f; v[k].z += 1e-9f; v[k].w += 1e-9f; base[at[k]] = v[k]; }
        }
    }
    if (mode & 2) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < slots; s += UN) {
            float4 v[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const int ss = min(s + k, slots - 1);
                const unsigned pick = cs ? (unsigned)__ldg(cs + (packs + pack) * slots + ss) : mix((unsigned)pack, (unsigned)ss, seedB) % (unsigned)vz;
                v[k] = base[((long long)ss * vz + pick) * 8];
            }
#pragma unroll
            for (int k = 0; k < UN; ++k)
                if (s + k < slots) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        }
        if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
    }
}


// the same passes with the rows staged in shared memory by 16-byte cp.async (LDGSTS): the rows in flight cost shared
// memory, not registers. Each lane owns UN slots of 16 B; issue UN copies, wait, consume.
template <int UN>
__global__ void __launch_bounds__(256) k_rows_async(float4 *w, long long packs, int slots, int vz, unsigned seedA, unsigned seedB, int mode, float *sink) {
    extern __shared__ float4 stage[];
    const int sub = threadIdx.x & 7;
    const long long pack = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
    if (pack >= packs) return;
    float4 *base = w + pack * slots * vz * 8 + sub;
    float4 *mine = stage + threadIdx.x;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(mine);
    for (int pass = 0; pass < 2; ++pass) {
        if (!((mode >> pass) & 1)) continue;
        const unsigned seed = pass ? seedB : seedA;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < slots; s += UN) {
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const int ss = min(s + k, slots - 1);
                const float4 *src = base + ((long long)ss * vz + mix((unsigned)pack, (unsigned)ss, seed) % (unsigned)vz) * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm0 + (unsigned)k * 256u * 16u), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                if (s + k >= slots) break;
                float4 v = mine[k * 256];
                if (pass == 0) {
                    v.x += 1e-9f; v.y += 1e-9f; v.z += 1e-9f; v.w += 1e-9f;
                    base[((long long)(s + k) * vz + mix((unsigned)pack, (unsigned)(s + k), seed) % (unsigned)vz) * 8] = v;
                } else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            }
        }
        if (pass == 1 && acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
    }
}

// one thread per (pair, visible cell): 16 threads = one pair. f: tiles of 16 x 32 floats (2 KB), r: rows of 16 floats
__global__ void __launch_bounds__(256) k_pairs(float *f, float *r, long long pairs, long long tiles, unsigned seed) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long pair = t >> 4;
    const int vzc = (int)(t & 15);
    if (pair >= pairs) return;
    const unsigned h = mix((unsigned)pair, (unsigned)(pair >> 32), seed);
    const long long tile = (long long)(((unsigned long long)h * (unsigned long long)tiles) >> 32);
    const int hz = (int)(mix(h, 7u, seed) & 31u);
    float *rp = r + (tile * 32 + hz) * 16 + vzc; // R: [tile][hz][vz]: the pair's 64-byte row
    float *fp = f + (tile * 16 + vzc) * 32 + hz; // F: [tile][vz][hz]: 16 floats 128 B apart
    const float d = 1e-9f;
    *rp = *rp + d;
    *fp = *fp + d;
}

int main(int argc, char **argv) {
    const long long packs = argc > 1 ? std::atoll(argv[1]) : 131072; // config 4 Predictor: 262144 columns, 2 per pack
    const int slots = 81, vz = 32;
    const long long rowsBytes = packs * slots * vz * 128;
    float4 *w = nullptr;
    float *sink = nullptr;
    CK(cudaMalloc(&w, (size_t)rowsBytes));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(w, 0, (size_t)rowsBytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const unsigned grid = (unsigned)((packs + 31) / 32);
    const char *names[4] = {"", "rmw", "gather", "rmw+gather"};
    for (int un = 0; un < 3; ++un)
        for (int mode = 1; mode <= 3; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 6; ++rep) {
                const unsigned sa = 1000u + 2u * rep, sb = 1001u + 2u * rep;
                CK(cudaEventRecord(e0));
                if (un == 0) k_rows<3><<<grid, 256>>>(w, packs, slots, vz, sa, sb, mode, sink, nullptr);
                else if (un == 1) k_rows<9><<<grid, 256>>>(w, packs, slots, vz, sa, sb, mode, sink, nullptr);
                else k_rows<27><<<grid, 256>>>(w, packs, slots, vz, sa, sb, mode, sink, nullptr);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep >= 2 && ms < best) best = ms;
            }
            const double bytes = (double)packs * slots * 128.0 * (mode == 1 ? 2 : mode == 2 ? 1 : 3);
            std::printf("{\"pattern\": \"rows\", \"pass\": \"%s\", \"rows_in_flight_per_lane\": %d, \"packs\": %lld, \"matrix_GB\": %.1f, \"bytes\": %.0f, \"ms\": %.4f, \"GBps\": %.1f}\n",
                        names[mode], un == 0 ? 3 : un == 1 ? 9 : 27, packs, rowsBytes / 1e9, bytes, best, bytes / best / 1e6);
        }
    // sensitivity of the fused pass (9 rows in flight per lane) to what the product kernel has and the plain pattern has not:
    // resident blocks per SM (the product runs 2 blocks of 256 threads: 110 registers) and a dependent index load
    {
        int *cs = nullptr;
        CK(cudaMalloc(&cs, (size_t)packs * slots * 2 * sizeof(int)));
        int *hcs = (int *)std::malloc((size_t)packs * slots * 2 * sizeof(int));
        unsigned x = 12345u;
        for (long long i = 0; i < packs * slots * 2; ++i) { x = x * 1664525u + 1013904223u; hcs[i] = (int)((x >> 16) % (unsigned)vz); }
        CK(cudaMemcpy(cs, hcs, (size_t)packs * slots * 2 * sizeof(int), cudaMemcpyHostToDevice));
        std::free(hcs);
        CK(cudaFuncSetAttribute(k_rows<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int occ[4] = {8, 4, 3, 2};
        for (int dep = 0; dep < 2; ++dep)
            for (int oi = 0; oi < 4; ++oi) {
                const size_t smem = occ[oi] == 8 ? 0 : (size_t)(220 * 1024 / occ[oi] - 2048);
                float best = 1e30f;
                for (int rep = 0; rep < 6; ++rep) {
                    CK(cudaEventRecord(e0));
                    k_rows<9><<<grid, 256, smem>>>(w, packs, slots, vz, 500u + rep, 900u + rep, 3, sink, dep ? cs : nullptr);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep >= 2 && ms < best) best = ms;
                }
                const double bytes = (double)packs * slots * 128.0 * 3;
                std::printf("{\"pattern\": \"rows\", \"pass\": \"rmw+gather\", \"rows_in_flight_per_lane\": 9, \"blocks_per_sm\": %d, \"dependent_index_load\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
                            occ[oi], dep, best, bytes / best / 1e6);
            }
        CK(cudaFree(cs));
    }
    // rows staged through shared memory (cp.async): rows in flight per lane 9 / 27 (blocks per SM follow from the shared memory)
    {
        CK(cudaFuncSetAttribute(k_rows_async<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_rows_async<27>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int un = 0; un < 2; ++un)
            for (int mode = 1; mode <= 3; ++mode) {
                float best = 1e30f;
                for (int rep = 0; rep < 6; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (un == 0) k_rows_async<9><<<grid, 256, 256 * 9 * 16>>>(w, packs, slots, vz, 300u + rep, 700u + rep, mode, sink);
                    else k_rows_async<27><<<grid, 256, 256 * 27 * 16>>>(w, packs, slots, vz, 300u + rep, 700u + rep, mode, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep >= 2 && ms < best) best = ms;
                }
                const double bytes = (double)packs * slots * 128.0 * (mode == 1 ? 2 : mode == 2 ? 1 : 3);
                std::printf("{\"pattern\": \"rows\", \"staging\": \"cp.async to shared\", \"pass\": \"%s\", \"rows_in_flight_per_lane\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
                            names[mode], un == 0 ? 9 : 27, best, bytes / best / 1e6);
            }
    }
    CK(cudaFree(w));

    // SparseCoder update: config 4 has 262144 x 81 tiles of 2 KB in F and as many rows x 32 in R; 736 k pairs per step
    const long long tiles = 262144LL * 81;
    float *f = nullptr, *r = nullptr;
    CK(cudaMalloc(&f, (size_t)tiles * 2048));
    CK(cudaMalloc(&r, (size_t)tiles * 2048));
    CK(cudaMemset(f, 0, (size_t)tiles * 2048));
    CK(cudaMemset(r, 0, (size_t)tiles * 2048));
    const long long pairCounts[3] = {736000, 1300000, 8350000};
    for (int pc = 0; pc < 3; ++pc) {
        const long long pairs = pairCounts[pc];
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0));
            k_pairs<<<(unsigned)((pairs * 16 + 255) / 256), 256>>>(f, r, pairs, tiles, 77u + rep);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep >= 2 && ms < best) best = ms;
        }
        std::printf("{\"pattern\": \"pairs\", \"pairs\": %lld, \"algorithmic_bytes\": %.0f, \"ms\": %.4f, \"ns_per_pair\": %.3f, \"algorithmic_GBps\": %.1f}\n", pairs,
                    pairs * 64.0, best, best * 1e6 / pairs, pairs * 64.0 / best / 1e6);
    }
    return 0;
}
