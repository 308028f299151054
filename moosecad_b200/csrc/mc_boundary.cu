// Mesh::boundary() (mesh/src/lib.rs:341-357) and the side-set loop (src/main.rs:87-106).
//
// The reference toggles every face of every tet in and out of a HashMap keyed by the sorted
// node triple.  Here only the faces that can possibly be unmatched are generated: a "full"
// cell (all five lattice tets emitted untouched) pairs all its internal faces by
// construction and its cell-face triangles pair with those of a full neighbour, so it only
// contributes the triangles facing a non-full neighbour or the lattice border; every other
// cell contributes all faces of its output tets.  The candidates (O(surface)) are radix
// sorted by key; a key that occurs an odd number of times is a boundary side, owned by its
// last occurrence in element order (the toggle semantics).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <limits>

#include "mc_internal.hpp"
#include "mc_kernels.cuh"

namespace mc {

// the two triangles of every cell face, as (tet, side) of the un-mirrored tile; derived once
// on the host from Tile::TETS / Tet4::face (a triangle lies in a cell face iff its three
// corners share a coordinate).
struct FaceTri { int8_t tet, side; };
static __constant__ FaceTri C_CELLFACE[6][2];  // [axis*2 + (coordinate==1)][k]

static void build_cellface_table(FaceTri out[6][2]) {
    static const int LAT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};
    static const int FACE[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    int cnt[6] = {0, 0, 0, 0, 0, 0};
    for (int t = 0; t < 5; ++t)
        for (int s = 0; s < 4; ++s)
            for (int ax = 0; ax < 3; ++ax)
                for (int v = 0; v < 2; ++v) {
                    bool all = true;
                    for (int k = 0; k < 3; ++k) all = all && LAT[TETS[t][FACE[s][k]]][ax] == v;
                    if (all) { out[ax * 2 + v][cnt[ax * 2 + v]++] = {(int8_t)t, (int8_t)s}; }
                }
}

// After the cut_lattice flip (nodes 0 and 1 swapped) the face table is applied to the flipped
// node order: side s of the emitted tet consists of flipped nodes FACE[s][*].  A triangle that
// is (un-flipped) side s0 keeps its node SET but may sit at a different side index.
__device__ __forceinline__ int side_after_flip(int s0) {
    // un-flipped nodes of side s0 -> positions after swapping 0<->1 -> which FACE row has that set
    int mask = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int n = C_FACE[s0][k];
        n = n < 2 ? 1 - n : n;
        mask |= 1 << n;
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int m = (1 << C_FACE[s][0]) | (1 << C_FACE[s][1]) | (1 << C_FACE[s][2]);
        if (m == mask) return s;
    }
    return -1;
}

Stuff make_stuff(const mc_mesh* m);  // mc_mesher.cu

__device__ __forceinline__ bool cell_is_full(const Stuff& S, int64_t cx, int64_t cy, int64_t cz) {
    const Lattice& L = S.L;
    if (cx < 0 || cy < 0 || cz < 0 || cx >= (int64_t)L.nx || cy >= (int64_t)L.ny || cz >= (int64_t)L.nz) return false;
    if (cz < (int64_t)S.cl0 || cz >= (int64_t)S.cl1) return false;
    const uint8_t cu = S.cuni[tile_of(S.g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz)];
    if (cu == CU_FULL) return true;
    if (cu == CU_EMPTY) return false;
    return (S.cinfo[cidx(S, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz)] & CI_FULL) != 0;
}

// number of candidate faces of a cell with `count` output tets
__device__ __forceinline__ uint32_t cell_face_count(const Stuff& S, uint32_t cx, uint32_t cy, uint32_t cz, bool full,
                                                    uint32_t count) {
    if (!full) return 4u * count;
    uint32_t n = 0;
    n += cell_is_full(S, (int64_t)cx - 1, cy, cz) ? 0u : 2u;
    n += cell_is_full(S, (int64_t)cx + 1, cy, cz) ? 0u : 2u;
    n += cell_is_full(S, cx, (int64_t)cy - 1, cz) ? 0u : 2u;
    n += cell_is_full(S, cx, (int64_t)cy + 1, cz) ? 0u : 2u;
    n += cell_is_full(S, cx, cy, (int64_t)cz - 1) ? 0u : 2u;
    n += cell_is_full(S, cx, cy, (int64_t)cz + 1) ? 0u : 2u;
    return n;
}

// (full?, output tets) of an owned cell of this tile
__device__ __forceinline__ void cell_state(const Stuff& S, uint8_t cu, uint32_t cx, uint32_t cy, uint32_t cz, bool& full,
                                           uint32_t& count, uint16_t& ci) {
    if (cu == CU_FULL) { full = true; count = 5; ci = CI_ALL_FULL; return; }
    ci = S.cinfo[cidx(S, cx, cy, cz)];
    full = (ci & CI_FULL) != 0;
    count = ci_count(ci);
}

static __global__ void __launch_bounds__(TILE_THREADS)
k_face_count(Stuff S, unsigned long long* __restrict__ tile_tot) {
    __shared__ unsigned long long s_acc;
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, blockIdx.x, X0, Y0, Z0);
    const uint8_t cu = S.cuni[blockIdx.x];
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    unsigned long long mine = 0;
    if (cu != CU_EMPTY && X0 < L.nx && Y0 < L.ny && zhi > zlo) {
        for (int r = 0; r < TILE_ROUNDS; ++r) {
            const uint32_t j = (uint32_t)r * TILE_THREADS + threadIdx.x;
            const uint32_t cx = X0 + (j & 15u), cy = Y0 + ((j >> 4) & 15u), cz = Z0 + (j >> 8);
            if (cx >= L.nx || cy >= L.ny || cz < zlo || cz >= zhi) continue;
            bool full; uint32_t count; uint16_t ci;
            cell_state(S, cu, cx, cy, cz, full, count, ci);
            if (count) mine += cell_face_count(S, cx, cy, cz, full, count);
        }
    }
    const unsigned long long tot = block_sum(mine, &s_acc);
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

__device__ __forceinline__ void sort3(unsigned long long& a, unsigned long long& b, unsigned long long& c) {
    unsigned long long t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
}

// emit candidate faces: key = sorted global node ids (k0 <= k1 <= k2), payload = tet * 4 + side
static __global__ void __launch_bounds__(TILE_THREADS)
k_face_emit(Stuff S, const unsigned long long* __restrict__ tile_fbase, unsigned long long tet_base,
            unsigned long long* __restrict__ key_hi, unsigned long long* __restrict__ key_lo,
            unsigned long long* __restrict__ payload) {
    __shared__ uint32_t s_scan[9];
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, blockIdx.x, X0, Y0, Z0);
    const uint8_t cu = S.cuni[blockIdx.x];
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    if (cu == CU_EMPTY || X0 >= L.nx || Y0 >= L.ny || zhi <= zlo) return;
    const uint32_t nxt = min(TS, L.nx - X0), nyt = min(TS, L.ny - Y0);
    const uint64_t tb = S.ttbase[blockIdx.x];
    uint64_t trun = tb, frun = tile_fbase[blockIdx.x];
    for (int r = 0; r < TILE_ROUNDS; ++r) {
        const uint32_t j = (uint32_t)r * TILE_THREADS + threadIdx.x;
        const uint32_t cx = X0 + (j & 15u), cy = Y0 + ((j >> 4) & 15u), cz = Z0 + (j >> 8);
        const bool owned = cx < L.nx && cy < L.ny && cz >= zlo && cz < zhi;
        bool full = false; uint32_t count = 0; uint16_t ci = 0;
        if (owned) cell_state(S, cu, cx, cy, cz, full, count, ci);
        const uint32_t nf = count ? cell_face_count(S, cx, cy, cz, full, count) : 0u;
        uint32_t ttot, ftot;
        const uint32_t tex = block_exscan(count, &ttot, s_scan);
        const uint32_t fex = block_exscan(nf, &ftot, s_scan);
        if (nf) {
            const Cell c = make_cell(cx, cy, cz);
            // slab-local index of the cell's first output tet (same numbering as k_tet_emit)
            uint64_t o = cu == CU_FULL ? tb + 5ull * (((cz - zlo) * nyt + (cy - Y0)) * nxt + (cx - X0)) : trun + tex;
            uint64_t f = frun + fex;
            if (full) {
                // tets 0..4 are emitted in order; only triangles on open cell faces are candidates
                for (int ax = 0; ax < 3; ++ax)
                    for (int v = 0; v < 2; ++v) {
                        // un-mirrored face (axis, v) sits at global offset v' = mirrored ? 1 - v : v
                        const bool mir = ax == 0 ? c.fx : (ax == 1 ? c.fy : c.fz);
                        const int gv = mir ? 1 - v : v;
                        int64_t ncx = cx, ncy = cy, ncz = cz;
                        (ax == 0 ? ncx : (ax == 1 ? ncy : ncz)) += gv ? 1 : -1;
                        if (cell_is_full(S, ncx, ncy, ncz)) continue;
                        for (int k = 0; k < 2; ++k) {
                            const FaceTri ft = C_CELLFACE[ax * 2 + v][k];
                            TetNodes tn;
                            load_tet(L, c, ft.tet, tn);
                            const int sd = c.flip() ? side_after_flip(ft.side) : ft.side;
                            unsigned long long g[3];
#pragma unroll
                            for (int q = 0; q < 3; ++q) g[q] = node_global_id<unsigned long long>(S, tn, C_FACE[sd][q]);
                            sort3(g[0], g[1], g[2]);
                            key_hi[f] = (g[0] << 32) | g[1];
                            key_lo[f] = g[2];
                            payload[f] = ((tet_base + o + (uint64_t)ft.tet) << 2) | (unsigned long long)sd;
                            ++f;
                        }
                    }
            } else {
                for (int t = 0; t < 5; ++t) {
                    TetNodes tn;
                    load_tet(L, c, t, tn);
                    TetOut to;
                    classify_tet<false>(L, c, t, tn, nullptr, (ci >> (CI_REMOVED_SHIFT + t)) & 1, to);
                    for (int q = 0; q < to.n; ++q) {
                        unsigned long long id[4];
#pragma unroll
                        for (int m = 0; m < 4; ++m) id[m] = node_global_id<unsigned long long>(S, tn, to.node(q, m));
                        for (int sd = 0; sd < 4; ++sd) {
                            unsigned long long g0 = id[C_FACE[sd][0]], g1 = id[C_FACE[sd][1]], g2 = id[C_FACE[sd][2]];
                            sort3(g0, g1, g2);
                            key_hi[f] = (g0 << 32) | g1;
                            key_lo[f] = g2;
                            payload[f] = ((tet_base + o) << 2) | (unsigned long long)sd;
                            ++f;
                        }
                        ++o;
                    }
                }
            }
        }
        trun += ttot;
        frun += ftot;
    }
}

static __global__ void k_iota(unsigned long long* p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
static __global__ void k_gather(const unsigned long long* __restrict__ src, const unsigned long long* __restrict__ idx,
                         uint64_t n, unsigned long long* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// one thread per sorted candidate: the first element of every run of equal keys counts the run
// and, when the length is odd, appends the run's last element (largest tet id) to the output
static __global__ void k_face_select(const unsigned long long* __restrict__ key_hi, const unsigned long long* __restrict__ key_lo,
                              const unsigned long long* __restrict__ payload,
                              const unsigned long long* __restrict__ order, uint64_t n,
                              unsigned long long* __restrict__ out_sides, unsigned long long* __restrict__ out_count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long me = order[i];
    const unsigned long long h = key_hi[me], l = key_lo[me];
    if (i > 0) {
        const unsigned long long pr = order[i - 1];
        if (key_hi[pr] == h && key_lo[pr] == l) return;  // not a run start
    }
    uint64_t j = i + 1;
    unsigned long long lastp = payload[me];
    while (j < n) {
        const unsigned long long nx = order[j];
        if (key_hi[nx] != h || key_lo[nx] != l) break;
        const unsigned long long p = payload[nx];
        if (p > lastp) lastp = p;
        ++j;
    }
    if ((j - i) & 1ull) {
        const unsigned long long at = atomicAdd(out_count, 1ull);
        out_sides[at] = lastp;  // packed elem * 4 + side; sorted and unpacked afterwards
    }
}

static __global__ void k_unpack_sides(const unsigned long long* __restrict__ packed, uint64_t n,
                                      unsigned long long* __restrict__ sides) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { sides[2 * i] = packed[i] >> 2; sides[2 * i + 1] = packed[i] & 3ull; }
}

// side-set membership: |approx_value(p, EPS)| <= EPS at the three nodes of every surface side
template <typename IndexT>
__global__ void k_sideset(const uint64_t* __restrict__ prog, const double* __restrict__ coords,
                          const IndexT* __restrict__ tets, int64_t vertex_base, unsigned long long tet_base,
                          const unsigned long long* __restrict__ sides, uint64_t n, uint8_t* __restrict__ member) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t ic = i < n ? i : n - 1;
    const unsigned long long e = sides[2 * ic] - tet_base, s = sides[2 * ic + 1];
    const double EPS = 2.220446049250313e-16;  // f64::EPSILON
    bool add = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t id = (int64_t)tets[4 * e + C_FACE[s][k]] - vertex_base;
        const double phi = vm_eval<true>(prog, coords[3 * id], coords[3 * id + 1], coords[3 * id + 2]);
        if (phi > EPS || phi < -EPS) add = false;
    }
    if (i < n) member[i] = add ? 1 : 0;
}

static int sort_u64(mc_ctx* c, unsigned long long* keys_in, unsigned long long* keys_out, unsigned long long* vals_in,
                    unsigned long long* vals_out, uint64_t n, int end_bit) {
    size_t tmp_bytes = 0;
    MC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int64_t)n, 0,
                                            end_bit, c->stream));
    DevBuf tmp;
    int rc = c->alloc(tmp_bytes, tmp);
    if (rc != MC_OK) return rc;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int64_t)n,
                                                    0, end_bit, c->stream);
    c->launches += 4;
    c->release(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    return MC_OK;
}

int compute_boundary(mc_mesh* m) {
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const Lattice& L = m->L;
    if (m->c0 != 0 || m->c1 != L.nz)
        return fail(MC_ERR_UNSUPPORTED, "Mesh::boundary on a z-slab: gather the slabs on one device first");
    {
        // per call: constant memory is per device, and the upload is a few hundred bytes
        FaceTri tab[6][2];
        build_cellface_table(tab);
        MC_CUDA(cudaMemcpyToSymbolAsync(C_CELLFACE, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, s));
    }
    cudaEvent_t e0, e1;
    MC_CUDA(cudaEventCreate(&e0));
    MC_CUDA(cudaEventCreate(&e1));
    MC_CUDA(cudaEventRecord(e0, s));
    DevBuf ftot, fbase, dtotal;
    int rc = c->alloc((m->n_tiles + 1) * 8, ftot);
    if (rc == MC_OK) rc = c->alloc((m->n_tiles + 1) * 8, fbase);
    if (rc == MC_OK) rc = c->alloc(4 * 8, dtotal);
    if (rc != MC_OK) return rc;
    MC_CUDA(cudaMemsetAsync(dtotal.p, 0, 32, s));
    unsigned long long* totals = (unsigned long long*)dtotal.p;
    uint64_t nf = 0;
    const Stuff S = make_stuff(m);
    if (m->n_tiles > 0) {
        k_face_count<<<(unsigned)m->n_tiles, TILE_THREADS, 0, s>>>(S, (unsigned long long*)ftot.p);
        c->launches += 1;
        MC_CUDA(cudaGetLastError());
        rc = scan_tiles(c, ftot.p, m->n_tiles, fbase.p, nullptr, m->d_scan_sums.p, totals);
        if (rc != MC_OK) return rc;
        unsigned long long h[2];
        MC_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, s));
        MC_CUDA(cudaStreamSynchronize(s));
        nf = h[0];  // face totals never exceed 32 bits per tile, the scan keeps 64-bit sums
    }
    m->n_sides = 0;
    m->h_sides.clear();
    if (nf > 0) {
        DevBuf khi, klo, pay, ia, ib, ka, kb;
        rc = c->alloc(nf * 8, khi);
        if (rc == MC_OK) rc = c->alloc(nf * 8, klo);
        if (rc == MC_OK) rc = c->alloc(nf * 8, pay);
        if (rc == MC_OK) rc = c->alloc(nf * 8, ia);
        if (rc == MC_OK) rc = c->alloc(nf * 8, ib);
        if (rc == MC_OK) rc = c->alloc(nf * 8, ka);
        if (rc == MC_OK) rc = c->alloc(nf * 8, kb);
        if (rc == MC_OK) rc = c->alloc(nf * 16, m->d_sides);
        if (rc != MC_OK) return rc;
        k_face_emit<<<(unsigned)m->n_tiles, TILE_THREADS, 0, s>>>(
            S, (const unsigned long long*)fbase.p, m->tet_base, (unsigned long long*)khi.p, (unsigned long long*)klo.p,
            (unsigned long long*)pay.p);
        const unsigned nb = (unsigned)((nf + 255) / 256);
        k_iota<<<nb, 256, 0, s>>>((unsigned long long*)ia.p, nf);
        c->launches += 2;
        MC_CUDA(cudaGetLastError());
        // LSD: by the low key first, then (stable) by the high key
        rc = sort_u64(c, (unsigned long long*)klo.p, (unsigned long long*)ka.p, (unsigned long long*)ia.p,
                      (unsigned long long*)ib.p, nf, 64);
        if (rc != MC_OK) return rc;
        k_gather<<<nb, 256, 0, s>>>((const unsigned long long*)khi.p, (const unsigned long long*)ib.p, nf,
                                    (unsigned long long*)kb.p);
        c->launches++;
        rc = sort_u64(c, (unsigned long long*)kb.p, (unsigned long long*)ka.p, (unsigned long long*)ib.p,
                      (unsigned long long*)ia.p, nf, 64);
        if (rc != MC_OK) return rc;
        k_face_select<<<nb, 256, 0, s>>>((const unsigned long long*)khi.p, (const unsigned long long*)klo.p,
                                         (const unsigned long long*)pay.p, (const unsigned long long*)ia.p, nf,
                                         (unsigned long long*)ka.p, totals + 2);
        c->launches++;
        MC_CUDA(cudaGetLastError());
        unsigned long long ns = 0;
        MC_CUDA(cudaMemcpyAsync(&ns, totals + 2, 8, cudaMemcpyDeviceToHost, s));
        MC_CUDA(cudaStreamSynchronize(s));
        m->n_sides = ns;
        if (ns) {
            // sorted by (elem, side): the reference's HashSet order is arbitrary
            size_t tmp_bytes = 0;
            MC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (unsigned long long*)ka.p, (unsigned long long*)kb.p,
                                                   (int64_t)ns, 0, 64, s));
            DevBuf tmp;
            rc = c->alloc(tmp_bytes, tmp);
            if (rc != MC_OK) return rc;
            MC_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, (unsigned long long*)ka.p, (unsigned long long*)kb.p,
                                                   (int64_t)ns, 0, 64, s));
            k_unpack_sides<<<(unsigned)((ns + 255) / 256), 256, 0, s>>>((const unsigned long long*)kb.p, ns,
                                                                       (unsigned long long*)m->d_sides.p);
            c->launches += 5;
            MC_CUDA(cudaGetLastError());
            MC_CUDA(cudaStreamSynchronize(s));
            c->release(tmp);
        }
        c->release(khi); c->release(klo); c->release(pay); c->release(ia); c->release(ib); c->release(ka); c->release(kb);
    }
    MC_CUDA(cudaEventRecord(e1, s));
    // host view, sorted (the reference's HashSet order is arbitrary)
    m->h_sides.resize(m->n_sides * 2);
    if (m->n_sides) MC_CUDA(cudaMemcpyAsync(m->h_sides.data(), m->d_sides.p, m->n_sides * 16, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    MC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    m->stage_ms[7] = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->release(ftot); c->release(fbase); c->release(dtotal);
    m->have_boundary = true;
    return MC_OK;
}

int compute_sideset(mc_mesh* m, const mc_tape* tape, int32_t root, std::vector<uint64_t>& out) {
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    out.clear();
    if (m->n_sides == 0) return MC_OK;
    Program p;
    std::string err;
    // boundary.approx_value(point, f64::EPSILON), src/main.rs:95
    int rc = tape->compile(root, std::numeric_limits<double>::epsilon(), p, err);
    if (rc != MC_OK) return fail(rc, err);
    DevBuf dp, dm;
    rc = upload_program(c, p, dp);
    if (rc != MC_OK) return rc;
    rc = c->alloc(m->n_sides, dm);
    if (rc != MC_OK) return rc;
    cudaEvent_t e0, e1;
    MC_CUDA(cudaEventCreate(&e0));
    MC_CUDA(cudaEventCreate(&e1));
    MC_CUDA(cudaEventRecord(e0, s));
    const unsigned nb = (unsigned)((m->n_sides + 255) / 256);
    if (m->index_bytes == 4)
        k_sideset<uint32_t><<<nb, 256, 0, s>>>((const uint64_t*)dp.p, (const double*)m->d_coords.p,
                                               (const uint32_t*)m->d_tets.p, (int64_t)m->vertex_base, m->tet_base,
                                               (const unsigned long long*)m->d_sides.p, m->n_sides, (uint8_t*)dm.p);
    else
        k_sideset<unsigned long long><<<nb, 256, 0, s>>>((const uint64_t*)dp.p, (const double*)m->d_coords.p,
                                                         (const unsigned long long*)m->d_tets.p, (int64_t)m->vertex_base,
                                                         m->tet_base, (const unsigned long long*)m->d_sides.p,
                                                         m->n_sides, (uint8_t*)dm.p);
    c->launches++;
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaEventRecord(e1, s));
    std::vector<uint8_t> member(m->n_sides);
    MC_CUDA(cudaMemcpyAsync(member.data(), dm.p, m->n_sides, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    MC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    m->stage_ms[8] += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (uint64_t i = 0; i < m->n_sides; ++i)
        if (member[i]) { out.push_back(m->h_sides[2 * i]); out.push_back(m->h_sides[2 * i + 1]); }
    c->release(dp);
    c->release(dm);
    return MC_OK;
}

}  // namespace mc
