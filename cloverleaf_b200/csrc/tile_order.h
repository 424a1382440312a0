// tile_order.h -- the order in which the persistent kernels' tile queue hands out tiles (host side, no CUDA):
// interior tiles first, rim tiles last, each part walked down bands of ~4096 columns, row by row inside a band.
//
// Tile (tx,ty) of a tw x th tiling owns the cells 1+tx*tw .. (tx+1)*tw x 1+ty*th .. (ty+1)*th and reads the columns
// j0-lo_x .. j0+tw-1+hi_x and rows k0-lo_y .. k0+th-1+hi_y (its TMA boxes).  It is INTERIOR if that neighbourhood lies
// entirely inside the cells 1..nx x 1..ny: no field of any kind has a halo cell there, so the tile depends on nothing
// a halo exchange / reflective boundary writes and may run next to it (common.cuh: programmatic dependent launch).
// Plain C++ so that the CPU test-suite can exercise it (tests/test_tile_order.py).
#pragma once
#include <vector>

namespace clv {

struct TileXY {
  int x, y;
};

inline bool tile_is_interior(int tx, int ty, int tw, int th, int lo_x, int hi_x, int lo_y, int hi_y, int nx, int ny) {
  const int j0 = 1 + tx * tw, k0 = 1 + ty * th;
  return j0 - lo_x >= 1 && j0 + tw - 1 + hi_x <= nx && k0 - lo_y >= 1 && k0 + th - 1 + hi_y <= ny;
}

// Fills `out` with all ntx*nty tiles; returns the number of interior tiles (they come first).
inline int build_tile_order(int ntx, int nty, int tw, int th, int lo_x, int hi_x, int lo_y, int hi_y, int nx, int ny,
                            std::vector<TileXY>& out) {
  constexpr int BAND_COLS = 4096;
  const int band_tiles = (ntx * tw <= BAND_COLS + 256) ? ntx : (BAND_COLS / tw > 0 ? BAND_COLS / tw : 1);
  out.clear();
  out.reserve((size_t)ntx * nty);
  int n_interior = 0;
  for (int pass = 0; pass < 2; ++pass) {  // 0: interior tiles, 1: rim tiles
    for (int x0 = 0; x0 < ntx; x0 += band_tiles) {
      const int w = (ntx - x0 < band_tiles) ? ntx - x0 : band_tiles;
      for (int ty = 0; ty < nty; ++ty)
        for (int tx = x0; tx < x0 + w; ++tx)
          if (tile_is_interior(tx, ty, tw, th, lo_x, hi_x, lo_y, hi_y, nx, ny) == (pass == 0)) out.push_back(TileXY{tx, ty});
    }
    if (pass == 0) n_interior = (int)out.size();
  }
  return n_interior;
}

}  // namespace clv
