// Commit-time tile classification, the part that needs no CUDA: from the per-tile summary (which info values occur, where, how
// often) to rectangles of one info value each.  A tile whose cells decompose into such rectangles is a UNIFORM tile (k_uniform never
// reads its per-cell info plane); everything else is GENERAL.  Header-only and free of CUDA types so that the CPU test suite can
// drive it with synthetic tiles (tests/test_tile_rectangles.py, tests/cpu/tile_rects_check.cpp).
#pragma once

#include <vector>

namespace chiml {

// per component: the (up to) TS_NV distinct non-zero info values of the tile in descending order, each with the bounding
// rectangle and count of its cells; the total count, and whether more than TS_NV values occur
constexpr int TS_NV = 6;
struct TileSummary
{
    unsigned info[3][TS_NV], rect[3][TS_NV], count[3][TS_NV];
    unsigned total[3], other[3];
    unsigned bytes; unsigned pad;
};

// rectangle packing: xlo | xhi << 8 | zlo << 16 | zhi << 24, tile-local, hi exclusive
inline unsigned rect_pack(unsigned x0, unsigned x1, unsigned z0, unsigned z1) { return x0 | (x1 << 8) | (z0 << 16) | (z1 << 24); }
inline unsigned rect_area(unsigned r) { return (((r >> 8) & 0xFF) - (r & 0xFF)) * ((r >> 24) - ((r >> 16) & 0xFF)); }

struct TileVal { unsigned info, rect; };

// Rectangles of component c of a tile.  Returns false when the cells do not decompose (more than TS_NV values, or a value whose
// cells are neither a rectangle nor a rectangle with one rectangular hole filled by another value).
//
// A value whose box has another value's rectangle cut out of it -- an object narrower than the tile, an object edge or corner inside
// the tile -- becomes the up to four rectangles around the hole: the full-width strips below and above it and the pieces left and
// right of it.  Why that is exact: all cells of the value lie in its bounding box and none in the hole (the hole's own value fills it
// completely: area == count); box minus hole has area(box) - area(hole) cells; if that equals the value's count, every one of them
// carries the value.
inline bool tile_rectangles(const TileSummary& ts, int c, std::vector<TileVal>& out)
{
    out.clear();
    if(!ts.total[c]) return true;
    if(ts.other[c]) return false;
    int nv = 0;
    while(nv < TS_NV && ts.count[c][nv]) ++nv;
    for(int w = 0; w < nv; ++w)
    {
        const unsigned rw = ts.rect[c][w];
        if(rect_area(rw) == ts.count[c][w]) { out.push_back({ts.info[c][w], rw}); continue; }
        bool split = false;
        const unsigned wx0 = rw & 0xFF, wx1 = (rw >> 8) & 0xFF, wz0 = (rw >> 16) & 0xFF, wz1 = rw >> 24;
        for(int h = 0; h < nv && !split; ++h)
        {
            if(h == w || rect_area(ts.rect[c][h]) != ts.count[c][h]) continue;
            const unsigned rh = ts.rect[c][h];
            const unsigned hx0 = rh & 0xFF, hx1 = (rh >> 8) & 0xFF, hz0 = (rh >> 16) & 0xFF, hz1 = rh >> 24;
            if(hx0 < wx0 || hx1 > wx1 || hz0 < wz0 || hz1 > wz1) continue;                    // not inside the box
            if(rect_area(rw) - rect_area(rh) != ts.count[c][w]) continue;                     // other values in the box too
            auto put = [&](unsigned x0, unsigned x1, unsigned z0, unsigned z1) { if(x1 > x0 && z1 > z0) out.push_back({ts.info[c][w], rect_pack(x0, x1, z0, z1)}); };
            put(wx0, wx1, wz0, hz0);     // strip below the hole (smaller z), full width
            put(wx0, wx1, hz1, wz1);     // strip above
            put(wx0, hx0, hz0, hz1);     // left of the hole
            put(hx1, wx1, hz0, hz1);     // right of the hole
            split = true;
        }
        if(!split) { out.clear(); return false; }
    }
    return true;
}

// How many rectangles of a component one record (one k_uniform block) carries: two when the tile is cut along z only (every warp,
// one z row, then lies in one rectangle and takes the column path); one when any cut runs along x, so that no warp has to run the
// two-rectangle body
inline int rectangles_per_record(const std::vector<TileVal> vals[3])
{
    for(int c = 0; c < 3; ++c)
        for(size_t w = 1; w < vals[c].size(); ++w)
            if((vals[c][w].rect & 0xFFFFu) != (vals[c][0].rect & 0xFFFFu)) return 1;
    return 2;
}

} // namespace chiml
