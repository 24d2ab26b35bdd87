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

// Rectangles of component c of a tile.  Returns false when the cells do not decompose (more than TS_NV values, a value whose cells
// are not "its bounding box minus rectangles of other values", or more than TILE_MAX_RECTS rectangles).
//
// A value whose box has other values' rectangles cut out of it -- an object narrower than the tile, object edges and corners, two
// objects in one tile -- is decomposed by a sweep along z: the z edges of the box and of the holes cut the box into bands; inside a
// band every hole either spans the band or misses it, so the band minus its holes is a list of x intervals; bands with the same
// interval list are merged.  Why that is exact: all cells of the value lie in its bounding box and none in a hole (a hole is a value
// that fills its own rectangle completely: area == count); box minus holes has area(box) - sum area(hole ∩ box) cells; if that
// equals the value's count, every one of them carries the value.
constexpr size_t TILE_MAX_RECTS = 12;
inline bool tile_rectangles(const TileSummary& ts, int c, std::vector<TileVal>& out)
{
    out.clear();
    if(!ts.total[c]) return true;
    if(ts.other[c]) return false;
    int nv = 0;
    while(nv < TS_NV && ts.count[c][nv]) ++nv;
    struct Box { unsigned x0, x1, z0, z1; };
    auto unpack = [](unsigned r) { return Box{r & 0xFF, (r >> 8) & 0xFF, (r >> 16) & 0xFF, r >> 24}; };
    for(int w = 0; w < nv; ++w)
    {
        const unsigned rw = ts.rect[c][w];
        if(rect_area(rw) == ts.count[c][w]) { out.push_back({ts.info[c][w], rw}); continue; }
        const Box B = unpack(rw);
        // holes: the parts inside the box of the other values that fill their own rectangles
        std::vector<Box> holes;
        unsigned holeArea = 0;
        for(int h = 0; h < nv; ++h)
        {
            if(h == w || rect_area(ts.rect[c][h]) != ts.count[c][h]) continue;
            Box H = unpack(ts.rect[c][h]);
            H.x0 = H.x0 > B.x0 ? H.x0 : B.x0; H.x1 = H.x1 < B.x1 ? H.x1 : B.x1;
            H.z0 = H.z0 > B.z0 ? H.z0 : B.z0; H.z1 = H.z1 < B.z1 ? H.z1 : B.z1;
            if(H.x0 >= H.x1 || H.z0 >= H.z1) continue;
            holes.push_back(H);
            holeArea += (H.x1 - H.x0) * (H.z1 - H.z0);
        }
        if(holes.empty() || rect_area(rw) - holeArea != ts.count[c][w]) { out.clear(); return false; }   // other values in the box too
        // z edges -> bands
        std::vector<unsigned> zs = {B.z0, B.z1};
        for(const Box& H : holes) { zs.push_back(H.z0); zs.push_back(H.z1); }
        for(size_t i = 0; i < zs.size(); ++i) for(size_t j = i + 1; j < zs.size(); ++j) if(zs[j] < zs[i]) { unsigned t = zs[i]; zs[i] = zs[j]; zs[j] = t; }
        std::vector<std::vector<unsigned>> bandIv;   // per band: x0, x1, x0, x1, ... of the value's intervals
        std::vector<unsigned> bandZ0, bandZ1;
        for(size_t i = 0; i + 1 < zs.size(); ++i)
        {
            const unsigned za = zs[i], zb = zs[i + 1];
            if(za == zb) continue;
            std::vector<Box> in;
            for(const Box& H : holes) if(H.z0 <= za && H.z1 >= zb) in.push_back(H);
            for(size_t a2 = 0; a2 < in.size(); ++a2) for(size_t b2 = a2 + 1; b2 < in.size(); ++b2) if(in[b2].x0 < in[a2].x0) { Box t = in[a2]; in[a2] = in[b2]; in[b2] = t; }
            std::vector<unsigned> iv;
            unsigned x = B.x0;
            for(const Box& H : in) { if(H.x0 > x) { iv.push_back(x); iv.push_back(H.x0); } if(H.x1 > x) x = H.x1; }
            if(B.x1 > x) { iv.push_back(x); iv.push_back(B.x1); }
            if(!bandIv.empty() && bandIv.back() == iv && bandZ1.back() == za) bandZ1.back() = zb;
            else { bandIv.push_back(iv); bandZ0.push_back(za); bandZ1.push_back(zb); }
        }
        for(size_t k = 0; k < bandIv.size(); ++k)
            for(size_t i = 0; i + 1 < bandIv[k].size(); i += 2)
                out.push_back({ts.info[c][w], rect_pack(bandIv[k][i], bandIv[k][i + 1], bandZ0[k], bandZ1[k])});
        if(out.size() > TILE_MAX_RECTS) { out.clear(); return false; }
    }
    if(out.size() > TILE_MAX_RECTS) { out.clear(); return false; }
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
