// Lattice topology of tessellate3d's 5-tet cubic tiling, order-free restatement for the GPU.
//
//   Tile::LATTICE / Tile::TETS                       tessellate3d/src/lib.rs:434-454
//   per-axis mirroring of odd cells + node 0<->1 flip tessellate3d/src/lib.rs:59-74,126-128
//   Tet4 edge / face tables                           mesh/src/lib.rs:650-653,661-668
//
// Consequences used throughout: the central tet's corners are the lattice
// points with even X+Y+Z; the edge set is the axis edges plus the face diagonals between even
// points; every undirected edge is owned by the endpoint from which its direction is one of the
// nine canonical directions below (last non-zero component positive).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mc {

struct Lattice {
    uint32_t nx, ny, nz;  // cells of the global lattice
    uint32_t NX, NY, NZ;  // points per axis (n + 1)
    uint32_t PX;          // row pitch of the per-point arrays in points: NX rounded up to a multiple of 16, so that
                          // every 16-point tile row of phi is one aligned 128-byte line and rows are 16-byte aligned (TMA)
    uint32_t pz0, pz1;    // stored point planes [pz0, pz1] of phi / wcode / vmask / vbase
    double ox, oy, oz;    // bbox.min (tessellate3d/src/lib.rs:57)
    double res;
    const double* __restrict__ phi;  // SDF at lattice points (pre-warp), x fastest, z slowest
    uint8_t* wcode;                  // bits 0..6 warp code (0 = not warped), bit 7 = zero-vertex used
    // Proven sign per 16^3 tile (+-1, 0 = not proven), or null.  The SDF kernel writes NOTHING into phi
    // for a proven tile: only the sign of such a point is ever needed (DESIGN.md "Points nobody
    // reads") and every reader that can meet one asks phi_at() / sign tables instead of the array.
    const int8_t* __restrict__ tsign;
    uint32_t tntx, tnty;             // tile grid (x, y) of tsign; z origin is pz0
};

// The lattice is processed in tiles of TS^3 points; tile (i,j,k) also names the TS^3 CELLS whose
// lowest corner lies in it.  Tiles are laid out x fastest over the stored planes.
constexpr uint32_t TS = 16;
struct TileGrid {
    uint32_t ntx, nty, ntz;  // tiles per axis
    uint32_t z0;             // first stored plane (= Lattice::pz0)
};
__device__ __forceinline__ uint32_t tile_of(const TileGrid& g, uint32_t X, uint32_t Y, uint32_t Z) {
    return (((Z - g.z0) / TS) * g.nty + Y / TS) * g.ntx + X / TS;
}
__device__ __forceinline__ void tile_origin(const TileGrid& g, uint64_t t, uint32_t& X0, uint32_t& Y0, uint32_t& Z0) {
    X0 = (uint32_t)(t % g.ntx) * TS;
    const uint64_t r = t / g.ntx;
    Y0 = (uint32_t)(r % g.nty) * TS;
    Z0 = g.z0 + (uint32_t)(r / g.nty) * TS;
}

__device__ __forceinline__ size_t pidx(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    return ((size_t)(Z - L.pz0) * L.NY + Y) * L.PX + X;
}
// SDF value at a lattice point as far as any consumer may look at it: the stored value, or +-1 in a
// tile whose sign is proven (nothing is stored there)
__device__ __forceinline__ int8_t tile_sign_at(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    return L.tsign ? L.tsign[(((Z - L.pz0) / TS) * L.tnty + Y / TS) * L.tntx + X / TS] : (int8_t)0;
}
__device__ __forceinline__ double phi_at(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    const int8_t s = tile_sign_at(L, X, Y, Z);
    return s ? (double)s : L.phi[pidx(L, X, Y, Z)];
}
// tessellate3d/src/lib.rs:84-103: (i as f64) * resolution + origin, per component
__device__ __forceinline__ double lat_x(const Lattice& L, uint32_t X) { return (double)X * L.res + L.ox; }
__device__ __forceinline__ double lat_y(const Lattice& L, uint32_t Y) { return (double)Y * L.res + L.oy; }
__device__ __forceinline__ double lat_z(const Lattice& L, uint32_t Z) { return (double)Z * L.res + L.oz; }

// Tile::LATTICE, Tile::TETS, Tet4::edge, Tet4::face
static __constant__ int8_t C_LAT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                   {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static __constant__ int8_t C_TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};
static __constant__ int8_t C_EDGE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static __constant__ int8_t C_FACE[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};

// The 18 neighbour directions of an even lattice point (first 6: axis, shared with odd points).
// Index d < 9 are the canonical (owning) directions; d + 9 is the opposite direction.
static __constant__ int8_t C_DIR[18][3] = {
    {1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},   {-1, 1, 0}, {1, 0, 1},  {-1, 0, 1}, {0, 1, 1},  {0, -1, 1},
    {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {-1, 0, -1}, {1, 0, -1}, {0, -1, -1}, {0, 1, -1}};

// direction (dx,dy,dz) in {-1,0,1}^3 -> index into C_DIR, or -1 (not a lattice edge direction)
// 27-entry table indexed by (dx+1) + 3*(dy+1) + 9*(dz+1)
static __constant__ int8_t C_DIR_INDEX[27] = {
    /* dz=-1 */ -1, 16, -1, 14, 11, 15, -1, 17, -1,
    /* dz= 0 */ 12, 10, 13, 9, -1, 0, 4, 1, 3,
    /* dz=+1 */ -1, 8, -1, 6, 2, 5, -1, 7, -1};
__device__ __forceinline__ int dir_index(int dx, int dy, int dz) {
    return C_DIR_INDEX[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)];
}

struct Cell {
    uint32_t cx, cy, cz;
    bool fx, fy, fz;  // mirrored along the axis iff the cell coordinate is odd
    __device__ __forceinline__ bool flip() const { return fx ^ fy ^ fz; }
};

__device__ __forceinline__ Cell make_cell(uint32_t cx, uint32_t cy, uint32_t cz) {
    Cell c;
    c.cx = cx; c.cy = cy; c.cz = cz;
    c.fx = cx & 1u; c.fy = cy & 1u; c.fz = cz & 1u;
    return c;
}

// global lattice coordinates of tile corner `k` of a cell
__device__ __forceinline__ void corner_xyz(const Cell& c, int k, uint32_t& X, uint32_t& Y, uint32_t& Z) {
    const int lx = C_LAT[k][0], ly = C_LAT[k][1], lz = C_LAT[k][2];
    X = c.cx + (uint32_t)(c.fx ? 1 - lx : lx);
    Y = c.cy + (uint32_t)(c.fy ? 1 - ly : ly);
    Z = c.cz + (uint32_t)(c.fz ? 1 - lz : lz);
}

// tile corner index of the lattice point (X,Y,Z) seen from cell c (the point must be a corner)
__device__ __forceinline__ int corner_of(const Cell& c, uint32_t X, uint32_t Y, uint32_t Z) {
    const int lx = (int)(X - c.cx), ly = (int)(Y - c.cy), lz = (int)(Z - c.cz);
    const int ux = c.fx ? 1 - lx : lx, uy = c.fy ? 1 - ly : ly, uz = c.fz ? 1 - lz : lz;
    // LATTICE index by (uz, uy, ux): z=0: 0 1 / 3 2 ; z=1: 4 5 / 7 6
    const int8_t T[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    return T[uz * 4 + uy * 2 + ux];
}

// tile corner of node `n` (0..3) of tet `t` AFTER the cut_lattice flip (nodes 0 and 1 swapped
// when fx^fy^fz, tessellate3d/src/lib.rs:126-128)
__device__ __forceinline__ int tet_corner(const Cell& c, int t, int n) {
    const int m = (c.flip() && n < 2) ? (1 - n) : n;
    return C_TETS[t][m];
}

}  // namespace mc
