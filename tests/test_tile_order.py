"""The tile order of the persistent kernels' ticket queue (cloverleaf_b200/csrc/tile_order.h, plain C++): interior
tiles -- whose TMA boxes contain no halo cell and therefore depend on no halo exchange -- first, rim tiles last.  A tile
wrongly classified interior would read halo cells while the exchange that writes them is still in flight (PDL), so
the classification is checked here against a brute-force statement of the rule, for the box shapes the kernels use."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = r'''
#include "tile_order.h"
extern "C" int tile_order_c(int ntx, int nty, int tw, int th, int lo_x, int hi_x, int lo_y, int hi_y, int nx, int ny, int* out) {
  std::vector<clv::TileXY> v;
  const int n = clv::build_tile_order(ntx, nty, tw, th, lo_x, hi_x, lo_y, hi_y, nx, ny, v);
  for (size_t i = 0; i < v.size(); ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; }
  return n;
}
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("tile_order")
    (d / "t.cpp").write_text(SRC)
    so = d / "libtile_order.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "cloverleaf_b200", "csrc"),
                    "-o", str(so), str(d / "t.cpp")], check=True)
    return ctypes.CDLL(str(so))


# (tile w, tile h, lo_x, hi_x, lo_y, hi_y, extra columns, extra rows of the tiled range) of the production launches
# (fuse.cu, advec_tma.cu): the vertex kernels tile 1..n+1, the y march of advec_cell tiles the rows 1..ny+2
SHAPES = {
    "timestep": (32, 16, 2, 2, 1, 1, 0, 0), "timestep_one_row": (32, 8, 2, 2, 1, 1, 0, 0), "pdv_predict": (32, 8, 0, 2, 0, 1, 0, 0), "lagrange_correct": (64, 8, 2, 2, 1, 1, 1, 1),
    "advec_cell_x": (60, 8, 2, 4, 0, 1, 0, 0), "advec_cell_y_three_phase": (32, 13, 2, 2, 2, 3, 0, 0),
    "advec_cell_y_march": (32, 32, 2, 2, 2, 3, 0, 2),
    "advec_mom_x": (60, 8, 2, 2, 1, 1, 1, 1), "advec_mom_y_three_phase": (32, 20, 2, 2, 2, 2, 1, 1),
    "advec_mom_y_march": (32, 24, 2, 2, 2, 2, 1, 1),
}


@pytest.mark.parametrize("kernel", sorted(SHAPES))
@pytest.mark.parametrize("nx,ny", [(3840, 3840), (1920, 960), (250, 130), (61, 37), (1, 1), (15360, 64)])
def test_interior_first_order(lib, kernel, nx, ny):
    tw, th, lo_x, hi_x, lo_y, hi_y, ex, ey = SHAPES[kernel]
    ntx, nty = (nx + ex + tw - 1) // tw, (ny + ey + th - 1) // th
    out = np.zeros(2 * ntx * nty, dtype=np.int32)
    n_int = lib.tile_order_c(ntx, nty, tw, th, lo_x, hi_x, lo_y, hi_y, nx, ny, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    tiles = [tuple(t) for t in out.reshape(-1, 2).tolist()]
    assert sorted(tiles) == [(x, y) for x in range(ntx) for y in range(nty)]  # a permutation of all tiles

    def box_inside(tx, ty):  # brute force: every cell the tile's boxes cover lies in 1..nx x 1..ny
        j0, k0 = 1 + tx * tw, 1 + ty * th
        return all(1 <= j <= nx for j in (j0 - lo_x, j0 + tw - 1 + hi_x)) and all(1 <= k <= ny for k in (k0 - lo_y, k0 + th - 1 + hi_y))

    assert all(box_inside(*t) for t in tiles[:n_int])
    assert not any(box_inside(*t) for t in tiles[n_int:])
    # the last tile row / column is never interior (its boxes reach past nx / ny), nor the first where the boxes
    # reach below cell 1 (pdv_predict's do not: it reads no low-side halo)
    for tx, ty in tiles[:n_int]:
        assert (tx < ntx - 1 or ntx == 1) and (ty < nty - 1 or nty == 1) or box_inside(tx, ty)
        assert (tx > 0 or lo_x == 0) and (ty > 0 or lo_y == 0)
    if (nx, ny) == (3840, 3840):
        assert n_int > 0.92 * len(tiles)  # big chunks: almost everything can overlap the exchange
    if (nx, ny) == (1920, 960):
        assert n_int > 0.80 * len(tiles)


def test_wide_chunks_are_walked_in_bands(lib):
    tw, th = 32, 8
    nx, ny = 15360, 256
    ntx, nty = nx // tw, ny // th
    out = np.zeros(2 * ntx * nty, dtype=np.int32)
    n_int = lib.tile_order_c(ntx, nty, tw, th, 2, 2, 1, 1, nx, ny, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    xs = out.reshape(-1, 2)[:n_int, 0]
    band = 4096 // tw
    # inside the interior part the band index never decreases: a band (4096 columns) is finished before the next starts
    assert np.all(np.diff(xs // band) >= 0)
