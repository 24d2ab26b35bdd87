// C ABI (include/chiml_gpu.h) of the B200 FDTD engine: setup from the reference's lists, commit
// (painting + pools), stepping, state access.  There is no CPU fallback: without a CUDA device
// chiml_gpu_create fails with CHIML_ERR_NO_DEVICE.
#include "chiml_kernels.cuh"

#include <cuda.h>          // CUtensorMap and the prototype of cuTensorMapEncodeTiled (resolved at run time: no -lcuda)
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>

using namespace chiml;

static thread_local std::string g_create_err;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if(_e != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                               \
            return CHIML_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while(0)

static int fail(ChimlCtx* ctx, int code, const std::string& msg) { ctx->err = msg; return code; }

static bool field_exists(const ChimlCtx* ctx, int f)
{
    const int c = f % 3;
    if(f >= CHIML_BX) return ctx->has_B && field_exists(ctx, CHIML_HX + c);       // B_[c] beside H_[c] (magnetic-dispersive media)
    const bool isH = f >= 3 && f < 6;
    if(f >= 6 && !ctx->g.has_D) return false;
    if(ctx->g.mode == CHIML_MODE_3D) return true;
    if(ctx->g.mode == CHIML_MODE_TE) return isH ? c == 2 : c != 2;
    return isH ? c != 2 : c == 2;
}

template <typename T> static int dev_alloc(ChimlCtx* ctx, T** p, size_t n, bool zero = true)
{
    if(n == 0) n = 1;
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    ctx->dev_bytes += n * sizeof(T);
    if(zero) CK(cudaMemsetAsync(*p, 0, n * sizeof(T), ctx->stream));
    return 0;
}
// buffers a neighbouring slab maps with CUDA IPC: whole multiples of 2 MiB, so that the driver never sub-allocates them
template <typename T> static int dev_alloc_ipc(ChimlCtx* ctx, T** p, size_t n)
{
    if(ctx->g.nranks <= 1) return dev_alloc(ctx, p, n);
    const size_t bytes = ((std::max<size_t>(n, 1) * sizeof(T) + IPC_GRANULE - 1) / IPC_GRANULE) * IPC_GRANULE;
    return dev_alloc(ctx, reinterpret_cast<char**>(p), bytes);
}
template <typename T> static int dev_upload(ChimlCtx* ctx, T** p, const std::vector<T>& v)
{
    int rc = dev_alloc(ctx, p, v.size(), false);
    if(rc) return rc;
    if(!v.empty()) CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// logical linear offset (+-1, +-lx, +-lx*lz) -> (dx, dy, dz)
static bool decode_offset(const ChimlCtx* ctx, long off, int d[3])
{
    d[0] = d[1] = d[2] = 0;
    if(off == 0) return true;
    const long a = off < 0 ? -off : off;
    const int s = off < 0 ? -1 : 1;
    if(a == 1) { d[0] = s; return true; }
    if(ctx->lz > 1 && a == ctx->lx) { d[2] = s; return true; }
    if(a == (long)ctx->lx * ctx->lz) { d[1] = s; return true; }
    return false;
}
static long phys_offset(const ChimlCtx* ctx, const int d[3]) { return d[0] + ctx->px * (d[2] + (long)ctx->lz * d[1]); }

extern "C" {

int chiml_gpu_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* chiml_gpu_last_error(const ChimlCtx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int chiml_gpu_create(const ChimlGridDesc* desc, int device, ChimlCtx** out)
{
    if(!desc || !out) { g_create_err = "null argument"; return CHIML_ERR_ARG; }
    *out = nullptr;
    if(desc->ln[0] < 3 || desc->ln[1] < 3 || desc->ln[2] < 1 || desc->mode < 0 || desc->mode > 2 || desc->n_objects < 1)
    { g_create_err = "bad grid description"; return CHIML_ERR_ARG; }
    if((desc->mode == CHIML_MODE_3D) != (desc->ln[2] > 1))
    { g_create_err = "mode / ln[2] mismatch: 3-D needs ln[2] > 1, 2-D needs ln[2] == 1"; return CHIML_ERR_ARG; }
    int ndev = chiml_gpu_device_count();
    if(ndev <= 0) { g_create_err = "no CUDA device visible: this engine has no CPU fallback"; return CHIML_ERR_NO_DEVICE; }
    if(device < 0 || device >= ndev) { g_create_err = "device index out of range"; return CHIML_ERR_ARG; }
    ChimlCtx* ctx = new ChimlCtx();
    ctx->g = *desc;
    ctx->device = device;
    ctx->lx = desc->ln[0]; ctx->ly = desc->ln[1]; ctx->lz = desc->ln[2];
    ctx->px = ((long)ctx->lx + 15) / 16 * 16;
    ctx->plane = ctx->px * ctx->lz;
    ctx->nphys = (size_t)ctx->plane * ctx->ly;
    ctx->nlogical = (size_t)ctx->lx * ctx->ly * ctx->lz;
    ctx->objs.resize(desc->n_objects);
    cudaError_t e = cudaSetDevice(device);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->hstream, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_push, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if(e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if(e != cudaSuccess) { g_create_err = std::string("CUDA init: ") + cudaGetErrorString(e); delete ctx; return CHIML_ERR_CUDA; }
    *out = ctx;
    return CHIML_OK;
}

void chiml_gpu_destroy(ChimlCtx* ctx)
{
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    if(ctx->stream) cudaStreamSynchronize(ctx->stream);
    if(ctx->imag) { ctx->imag->stream = nullptr; ctx->imag->real_part = nullptr; ctx->imag = nullptr; }      // the imaginary part ran on this context's stream
    if(ctx->real_part) { ctx->real_part->imag = nullptr; ctx->real_part = nullptr; }
    for(auto& p : ctx->d_field_base) cudaFree(p);
    for(int c = 0; c < 6; ++c)
    {
        cudaFree(ctx->d_info[c]); cudaFree(ctx->d_cls[c]); cudaFree(ctx->d_pf[c]);
        for(int k = 0; k < 2; ++k)
        {
            PmlPartDev& pp = ctx->pml[c][k];
            cudaFree(pp.d_F); cudaFree(pp.d_b); cudaFree(pp.d_c); cudaFree(pp.d_cmap); cudaFree(pp.d_psi);
        }
    }
    for(int c = 0; c < 6; ++c)
    {
        cudaFree(ctx->span[c].d_xmin); cudaFree(ctx->span[c].d_xmax); cudaFree(ctx->span[c].d_base); cudaFree(ctx->span[c].d_rows);
        for(int p = 0; p < MAX_POLES; ++p)
            for(int k = 0; k < 2; ++k) { cudaFree(ctx->d_P[c][p][k]); if(c < 3) cudaFree(ctx->d_oP[c][p][k]); }
        if(c < 3) for(int p = 0; p < MAX_POLES; ++p) cudaFree(ctx->d_dipg[c][p]);
    }
    for(int c = 0; c < 6; ++c)
    {
        cudaFree(ctx->d_prev_base[c]);
        for(int p = 0; p < MAX_CHI; ++p) for(int k = 0; k < 2; ++k) cudaFree(ctx->d_chi[c][p][k]);
    }
    cudaFree(ctx->d_prev_rows);
    cudaFree(ctx->d_info_node); cudaFree(ctx->d_cls_node);
    cudaFree(ctx->span_node.d_xmin); cudaFree(ctx->span_node.d_xmax); cudaFree(ctx->span_node.d_base); cudaFree(ctx->span_node.d_rows);
    cudaFree(ctx->d_src_amp);
    for(auto& f : ctx->d_tiles) for(auto& p : f) cudaFree(p);
    for(auto& d : ctx->detectors) cudaFree(d.d_ring);
    for(auto& d : ctx->dfts) { cudaFree(d.d_lines); cudaFree(d.d_re); cudaFree(d.d_im); }
    cudaFree(ctx->d_tw);
    for(auto& em : ctx->emitters)
    {
        cudaFree(em.d_h0); cudaFree(em.d_mu); cudaFree(em.d_gam_val); cudaFree(em.d_eps); cudaFree(em.d_gam_ptr); cudaFree(em.d_gam_col); cudaFree(em.d_loc);
        for(auto& p : em.d_P) cudaFree(p);
        cudaFree(em.d_rho);
        for(auto& p : em.d_f) cudaFree(p);
        cudaFree(em.d_pop_partial); cudaFree(em.d_pop);
    }
    for(auto& v : ctx->ev_pending) for(auto& pr : v) { cudaEventDestroy(pr[0]); cudaEventDestroy(pr[1]); }
    for(auto& e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->hstream);
    for(HaloPeer* pr : {&ctx->lower, &ctx->upper}) for(void* b : pr->opened) cudaIpcCloseMemHandle(b);
    cudaFree(ctx->d_flags); cudaFree(ctx->d_push_counter); cudaFree(ctx->d_persist_sa); cudaFree(ctx->d_tmaps); cudaFree(ctx->d_tfsf_incd);
    for(auto& t : ctx->tfsf) { cudaFree(t.d_pairs_D); cudaFree(t.d_pairs_U); cudaFree(t.d_ep_mu); }
    for(auto& g : ctx->d_oPy_ghost) cudaFree(g);
    cudaEventDestroy(ctx->ev_main); cudaEventDestroy(ctx->ev_push);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->hstream);
    if(ctx->stream && !ctx->is_imag_part) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int chiml_gpu_set_update_list(ChimlCtx* ctx, int kind, int comp, const ChimlRun* runs, size_t n)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_update_list after commit");
    if(kind < 0 || kind > 5 || comp < 0 || comp > 5 || (n && !runs)) return fail(ctx, CHIML_ERR_ARG, "set_update_list: bad kind/comp");
    if(kind == CHIML_LIST_CHID && n && (ctx->g.mode != CHIML_MODE_3D || !ctx->g.has_D || !ctx->has_B || ctx->g.nranks > 1))
        return fail(ctx, CHIML_ERR_UNSUPPORTED, "chiral lists need a 3-D grid with D and B grids (chiml_gpu_set_magnetic first) on a single slab");
    if(kind != CHIML_LIST_U && comp > 2 && n && !(ctx->has_B && (kind == CHIML_LIST_D || kind == CHIML_LIST_LORD || kind == CHIML_LIST_CHID)))
        return fail(ctx, CHIML_ERR_UNSUPPORTED, "magnetic lists need chiml_gpu_set_magnetic(has_B = 1) first (upB_ / upLorB_); magnetic oriented-dipole lists are outside the covered hot path");
    const long ncell = (long)ctx->nlogical;
    for(size_t i = 0; i < n; ++i)
    {
        const ChimlRun& r = runs[i];
        if(r.n < 1 || r.ind < 0 || (long)r.ind + r.n > ncell || r.obj < 0 || r.obj >= ctx->g.n_objects)
            return fail(ctx, CHIML_ERR_ARG, "set_update_list: run outside the grid or bad object index");
        if((r.ind % ctx->lx) + r.n > ctx->lx) return fail(ctx, CHIML_ERR_ARG, "set_update_list: run crosses a row end");
    }
    ctx->lists[kind][comp].runs.assign(runs, runs + n);
    return CHIML_OK;
}

int chiml_gpu_set_object(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, int use_or_dip, const double* dip)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_object after commit");
    if(obj < 0 || obj >= ctx->g.n_objects || npoles < 0) return fail(ctx, CHIML_ERR_ARG, "set_object: bad index");
    if(npoles > MAX_POLES) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_object: more than 12 poles per object");
    if(npoles > 0 && (!alpha || !xi || !gamma)) return fail(ctx, CHIML_ERR_ARG, "set_object: null pole constants");
    HostObj& o = ctx->objs[obj];
    o.npoles = npoles; o.use_or_dip = use_or_dip;
    o.alpha.assign(alpha, alpha + npoles); o.xi.assign(xi, xi + npoles); o.gamma.assign(gamma, gamma + npoles);
    o.dip.assign(3 * (size_t)npoles, 0.0);
    if(dip) o.dip.assign(dip, dip + 3 * (size_t)npoles);
    return CHIML_OK;
}

int chiml_gpu_set_dip_grid(ChimlCtx* ctx, int comp, int pole, const double* grid)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_dip_grid after commit");
    if(!grid) return fail(ctx, CHIML_ERR_ARG, "set_dip_grid: null grid");
    if(comp < 0 || comp > 2 || pole < 0) return fail(ctx, CHIML_ERR_ARG, "set_dip_grid: bad comp/pole");
    if(pole >= MAX_POLES) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_dip_grid: more than 12 poles per object");
    if(!field_exists(ctx, comp)) return fail(ctx, CHIML_ERR_ARG, "set_dip_grid: the field component does not exist in this mode");
    ctx->h_dipg[comp][pole] = grid;
    ctx->has_dipg = true;
    return CHIML_OK;
}

int chiml_gpu_set_ordip_pole_count(ChimlCtx* ctx, int n_poles_global)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_ordip_pole_count after commit");
    if(n_poles_global < 0) return fail(ctx, CHIML_ERR_ARG, "set_ordip_pole_count: negative count");
    if(n_poles_global > MAX_POLES) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_ordip_pole_count: more than 12 poles per object");
    ctx->nordip_global = n_poles_global;
    return CHIML_OK;
}

int chiml_gpu_set_object_chiral(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, const double* gamma_prev)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_object_chiral after commit");
    if(obj < 0 || obj >= (int)ctx->objs.size()) return fail(ctx, CHIML_ERR_ARG, "set_object_chiral: object index out of range");
    if(npoles < 0 || npoles > MAX_CHI) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_object_chiral: more chiral poles than MAX_CHI");
    if(npoles > 0 && (!alpha || !xi || !gamma || !gamma_prev)) return fail(ctx, CHIML_ERR_ARG, "set_object_chiral: NULL constant array");
    HostObj& o = ctx->objs[obj];
    o.nchi = npoles;
    o.calpha.assign(alpha, alpha + npoles); o.cxi.assign(xi, xi + npoles); o.cgamma.assign(gamma, gamma + npoles); o.cgprev.assign(gamma_prev, gamma_prev + npoles);
    return 0;
}

int chiml_gpu_set_prev_copy(ChimlCtx* ctx, const int32_t* rows, size_t nrows)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_prev_copy after commit");
    if(nrows && !rows) return fail(ctx, CHIML_ERR_ARG, "set_prev_copy: NULL rows");
    for(size_t q = 0; q < nrows; ++q)
    {
        const int32_t* b = rows + 4 * q;
        if(b[0] < 0 || b[1] < 0 || b[1] + b[0] > ctx->lx || b[2] < 0 || b[2] >= ctx->ly || b[3] < 0 || b[3] >= ctx->lz)
            return fail(ctx, CHIML_ERR_ARG, "set_prev_copy: a row leaves the grid");
    }
    ctx->h_prev_rows.assign(rows, rows + 4 * nrows);
    return 0;
}

int chiml_gpu_set_magnetic(ChimlCtx* ctx, int has_B, int pml_on_B)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_magnetic after commit");
    if(pml_on_B && !has_B) return fail(ctx, CHIML_ERR_ARG, "set_magnetic: the CPML cannot act on B without B grids");
    ctx->has_B = has_B ? 1 : 0; ctx->pml_on_B = pml_on_B ? 1 : 0;
    return 0;
}

int chiml_gpu_set_object_magnetic(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_object_magnetic after commit");
    if(obj < 0 || obj >= (int)ctx->objs.size()) return fail(ctx, CHIML_ERR_ARG, "set_object_magnetic: object index out of range");
    if(npoles < 0 || npoles > MAX_POLES) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_object_magnetic: more poles than MAX_POLES");
    if(npoles > 0 && (!alpha || !xi || !gamma)) return fail(ctx, CHIML_ERR_ARG, "set_object_magnetic: NULL constant array");
    HostObj& o = ctx->objs[obj];
    o.nmag = npoles;
    o.malpha.assign(alpha, alpha + npoles); o.mxi.assign(xi, xi + npoles); o.mgamma.assign(gamma, gamma + npoles);
    return 0;
}

int chiml_gpu_add_tfsf_surface(ChimlCtx* ctx, const ChimlTfsfSurface* t)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "add_tfsf_surface after commit");
    if(!t || t->comp < 0 || t->comp > 5 || !field_exists(ctx, t->comp)) return fail(ctx, CHIML_ERR_ARG, "add_tfsf_surface: target component 0..5 of this mode");
    if(ctx->g.nranks > 1) return fail(ctx, CHIML_ERR_UNSUPPORTED, "TFSF surfaces are covered for single-slab runs only");
    if((t->n < 0 && (t->npairs_D > 0 || t->npairs_U > 0)) || t->npairs_D < 0 || t->npairs_U < 0 || t->incd_len < 1 || t->incd_offset < 0 || t->stride_main < 1 ||
       (t->npairs_D > 0 && !t->pairs_D) || (t->npairs_U > 0 && !t->pairs_U))
        return fail(ctx, CHIML_ERR_ARG, "add_tfsf_surface: inconsistent surface record");
    if(t->npairs_D > 0 && (t->comp > 2 || !ctx->g.has_D)) return fail(ctx, CHIML_ERR_UNSUPPORTED, "add_tfsf_surface: D pairs need an E component and D grids (B targets are outside the covered hot path)");
    // every incident index a pair reaches must lie inside its line
    const long span = (long)(t->n - 1) * std::abs(t->stride_incd);
    for(int side = 0; side < 2; ++side)
    {
        const int32_t* pr = side ? t->pairs_U : t->pairs_D;
        for(int l = 0; l < (side ? t->npairs_U : t->npairs_D); ++l)
            if(t->n > 0 && (pr[2 * l] < 0 || pr[2 * l] + span >= t->incd_len)) return fail(ctx, CHIML_ERR_ARG, "add_tfsf_surface: a pair reads outside its incident line");
    }
    TfsfDev d;
    d.s = *t;
    d.h_pairs_D.assign(t->pairs_D, t->pairs_D + 2 * (size_t)t->npairs_D);
    d.h_pairs_U.assign(t->pairs_U, t->pairs_U + 2 * (size_t)t->npairs_U);
    if(t->ep_mu) d.h_ep_mu.assign(t->ep_mu, t->ep_mu + t->incd_len);
    d.s.pairs_D = nullptr; d.s.pairs_U = nullptr; d.s.ep_mu = nullptr;
    ctx->tfsf.push_back(std::move(d));
    return 0;
}

int chiml_gpu_set_periodic(ChimlCtx* ctx, int comp, const ChimlWrap* w)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_periodic after commit");
    if(comp < 0 || comp > 5 || !w) return fail(ctx, CHIML_ERR_ARG, "set_periodic: component 0..5 and a wrap description");
    if(!field_exists(ctx, comp)) return fail(ctx, CHIML_ERR_ARG, "set_periodic: the component does not exist in this mode");
    const bool twoD = ctx->lz == 1;
    // a slab of several takes the x / z wraps of its owned rows only (ymax = ny = -1): the y direction is the ring of ghost-row pushes
    const bool slab = ctx->g.nranks > 1;
    if(slab != (w->ymax < 0)) return fail(ctx, CHIML_ERR_ARG, slab ? "set_periodic: a slab of several takes ymax = ny = -1 (x / z wraps only; y is the slab ring)"
                                                                   : "set_periodic: ymax = -1 is for slabs of several");
    // the images must lie inside the arrays and the box must have an inside
    if(w->xmax < 2 || (!slab && (w->ymax < 2 || w->ymax > ctx->ly - 1 || w->ny != w->ymax)) || w->xmax > ctx->lx - 1 || w->nx != w->xmax - 1 ||
       (twoD ? (w->zmin != 0) : (w->zmin != 1 || w->zmax < 2 || w->zmax > ctx->lz - 1 || w->nz != w->zmax - 1)))
        return fail(ctx, CHIML_ERR_ARG, "set_periodic: the wrap box does not fit the grid (expected the arguments of applBCE_/applBCH_)");
    ctx->wrap[comp] = *w; ctx->has_wrap[comp] = true; ctx->periodic = true; ctx->ring = slab;
    return 0;
}

int chiml_gpu_set_persistent(ChimlCtx* ctx, int on)
{
    if(!ctx) return CHIML_ERR_ARG;
    ctx->persist_mode = on ? 1 : 0;
    return CHIML_OK;
}

int chiml_gpu_set_march(ChimlCtx* ctx, int fast_planes, int uniform_planes)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_march after commit");
    if(fast_planes < 0 || uniform_planes < 0) return fail(ctx, CHIML_ERR_ARG, "set_march: negative column length");
    ctx->march_fast = fast_planes; ctx->march_uniform = uniform_planes;
    return CHIML_OK;
}

int chiml_gpu_set_cpml(ChimlCtx* ctx, int comp, int part, int has_psi, const ChimlPsiParams* psi, size_t npsi, const ChimlGridParams* grid, size_t ngrid)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "set_cpml after commit");
    if(comp < 0 || comp > 5 || part < 0 || part > 1) return fail(ctx, CHIML_ERR_ARG, "set_cpml: bad comp/part");
    if((npsi && !psi) || (ngrid && !grid)) return fail(ctx, CHIML_ERR_ARG, "set_cpml: null list");
    HostPml& h = ctx->hpml[comp][part];
    h.present = 1; h.has_psi = has_psi;
    h.psi.assign(psi, psi + npsi);
    h.grid.assign(grid, grid + ngrid);
    return CHIML_OK;
}

int chiml_gpu_add_source(ChimlCtx* ctx, int field, const int32_t loc[3], const int32_t sz[3], int* slot)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "add_source after commit");
    if(!loc || !sz) return fail(ctx, CHIML_ERR_ARG, "add_source: null box");
    if(field < 0 || field >= 6 || !field_exists(ctx, field)) return fail(ctx, CHIML_ERR_ARG, "add_source: field absent in this mode");
    if((int)ctx->sources.size() >= MAX_SOURCES) return fail(ctx, CHIML_ERR_UNSUPPORTED, "too many sources");
    SourceDev s; s.field = field;
    const int ln[3] = {ctx->lx, ctx->ly, ctx->lz};
    for(int k = 0; k < 3; ++k)
    {
        if(sz[k] < 1 || loc[k] < 0 || loc[k] + sz[k] > ln[k]) return fail(ctx, CHIML_ERR_ARG, "add_source: box outside the local grid");
        s.loc[k] = loc[k]; s.sz[k] = sz[k];
    }
    if(slot) *slot = (int)ctx->sources.size();
    ctx->sources.push_back(s);
    return CHIML_OK;
}

int chiml_gpu_add_detector(ChimlCtx* ctx, int field, const int32_t loc[3], const int32_t sz[3], int every, int* slot)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "add_detector after commit");
    if(!loc || !sz) return fail(ctx, CHIML_ERR_ARG, "add_detector: null box");
    if(field < 0 || field >= CHIML_NFIELDS || !field_exists(ctx, field)) return fail(ctx, CHIML_ERR_ARG, "add_detector: field absent in this mode");
    if(every < 1) return fail(ctx, CHIML_ERR_ARG, "add_detector: interval must be >= 1 step");
    DetectorDev d; d.field = field; d.every = every;
    const int ln[3] = {ctx->lx, ctx->ly, ctx->lz};
    for(int k = 0; k < 3; ++k)
    {
        if(sz[k] < 1 || loc[k] < 0 || loc[k] + sz[k] > ln[k]) return fail(ctx, CHIML_ERR_ARG, "add_detector: box outside the local grid");
        d.loc[k] = loc[k]; d.sz[k] = sz[k];
    }
    d.sample_len = (size_t)sz[0] * sz[1] * sz[2];
    if(slot) *slot = (int)ctx->detectors.size();
    ctx->detectors.push_back(d);
    return CHIML_OK;
}

int chiml_gpu_add_dft(ChimlCtx* ctx, int field, int group, int every, int nfreq, int npts, int stride, const ChimlDftLine* lines, size_t nlines, size_t acc_len, int* slot)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "add_dft after commit");
    if(field < 0 || field >= CHIML_NFIELDS || !field_exists(ctx, field)) return fail(ctx, CHIML_ERR_ARG, "add_dft: field absent in this mode");
    if(group < 0 || group > 255 || every < 1 || nfreq < 1 || npts < 1 || (nlines && !lines)) return fail(ctx, CHIML_ERR_ARG, "add_dft: bad argument");
    if((int)ctx->dft_group_nfreq.size() <= group) ctx->dft_group_nfreq.resize(group + 1, 0);
    if(ctx->dft_group_nfreq[group] != 0 && ctx->dft_group_nfreq[group] != nfreq) return fail(ctx, CHIML_ERR_ARG, "add_dft: the sets of one group must share the frequency list");
    ctx->dft_group_nfreq[group] = nfreq;
    for(size_t l = 0; l < nlines; ++l)
    {
        const long last = (long)lines[l].ind + (long)(npts - 1) * stride;
        if(lines[l].ind < 0 || last >= (long)ctx->nlogical || last < 0) return fail(ctx, CHIML_ERR_ARG, "add_dft: line leaves the grid");
        if(lines[l].out < 0 || (size_t)lines[l].out + (size_t)nfreq * npts > acc_len) return fail(ctx, CHIML_ERR_ARG, "add_dft: line leaves the accumulator");
    }
    DftDev d;
    d.field = field; d.group = group; d.every = every; d.nfreq = nfreq; d.npts = npts; d.stride = stride; d.acc_len = acc_len;
    // fInGridInds_ is sized for the whole surface; a rank that holds part of it leaves the tail (0, 0).  Those entries would all add
    // the (zero) corner ghost cell into accumulator 0 -- harmless on the CPU, a read-modify-write race with the real line 0 here
    for(size_t l = 0; l < nlines; ++l)
        if(l == 0 || lines[l].ind != 0 || lines[l].out != 0) d.h_lines.push_back(lines[l]);
    d.nlines = d.h_lines.size();
    if(slot) *slot = (int)ctx->dfts.size();
    ctx->dfts.push_back(std::move(d));
    return CHIML_OK;
}

int chiml_gpu_add_emitters(ChimlCtx* ctx, const ChimlEmitterDesc* d, int* slot)
{
    if(!ctx || !d) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "add_emitters after commit");
    if(d->nlevel < 2 || d->nlevel > 6) return fail(ctx, CHIML_ERR_UNSUPPORTED, "add_emitters: 2 <= nlevel <= 6 supported");
    if(d->nsys < 1 || d->nemit < 0 || d->npop < 0 || d->npop > EMIT_MAX_POP || d->pop_every < 1)
        return fail(ctx, CHIML_ERR_ARG, "add_emitters: bad counts (at most 8 population detectors per object)");
    if(!d->h0 || !d->weight || !d->mu || !d->gam_ptr || !d->eps || (d->nemit && !d->loc) || (d->npop && !d->pop_level))
        return fail(ctx, CHIML_ERR_ARG, "add_emitters: null array");
    if(!ctx->g.has_D) return fail(ctx, CHIML_ERR_ARG, "add_emitters: emitter objects are D-cells, has_D must be set");
    EmitterDev em;
    em.d = *d;
    const int n2 = d->nlevel * d->nlevel;
    em.n2 = n2;
    const bool threeD = ctx->lz > 1;
    em.pz = threeD ? d->box_n[2] + 2 : 2;
    em.pbox = (size_t)(d->box_n[0] + 2) * (size_t)(d->box_n[1] + 2) * (size_t)em.pz;
    const int ln[3] = {ctx->lx, ctx->ly, ctx->lz};
    for(int k = 0; k < 3; ++k)
    {
        if(!threeD && k == 2) continue;
        // a slab whose last owned row lies just below the object holds a set without emitters and without box rows (box_n[1] == 0):
        // its E_y in that row still feels P_y of the object's first row, which the slab above pushes into this set's rim
        const int nmin = (k == 1 && d->nemit == 0 && ctx->g.nranks > 1) ? 0 : 1;
        if(d->box_n[k] < nmin || d->box_lo[k] < 0 || d->box_lo[k] + d->box_n[k] + 2 > ln[k])
            return fail(ctx, CHIML_ERR_ARG, "add_emitters: emitter box (plus its one-node rim) leaves the local grid");
    }
    for(int e = 0; e < d->nemit; ++e)
        for(int k = 0; k < 3; ++k)
            if(d->loc[3 * e + k] < 0 || d->loc[3 * e + k] >= (k == 2 && !threeD ? 1 : d->box_n[k]))
                return fail(ctx, CHIML_ERR_ARG, "add_emitters: emitter outside its box");
    const int nnz = d->gam_ptr[n2];
    for(int k = 0; k < nnz; ++k) if(d->gam_col[k] < 0 || d->gam_col[k] >= n2) return fail(ctx, CHIML_ERR_ARG, "add_emitters: gam column out of range");
    for(int p = 0; p < d->npop; ++p) if(d->pop_level[p] < 0 || d->pop_level[p] >= n2) return fail(ctx, CHIML_ERR_ARG, "add_emitters: population level out of range");
    em.h_h0.assign(d->h0, d->h0 + (size_t)d->nsys * n2 * 2);
    em.h_weight.assign(d->weight, d->weight + d->nsys);
    em.h_mu.assign(d->mu, d->mu + (size_t)3 * n2 * 2);
    em.h_gam_ptr.assign(d->gam_ptr, d->gam_ptr + n2 + 1);
    if(nnz) { em.h_gam_col.assign(d->gam_col, d->gam_col + nnz); em.h_gam_val.assign(d->gam_val, d->gam_val + nnz); }
    if(d->nemit) em.h_loc.assign(d->loc, d->loc + (size_t)3 * d->nemit);
    em.h_eps.assign(d->eps, d->eps + em.pbox);
    if(d->npop) em.h_pop_level.assign(d->pop_level, d->pop_level + d->npop);
    for(int c = 0; c < 3; ++c)
        for(int k = 0; k < 2 * n2; ++k) if(em.h_mu[(size_t)c * n2 * 2 + k] != 0.0) em.mu_present[c] = 1;
    if(slot) *slot = (int)ctx->emitters.size();
    ctx->emitters.push_back(std::move(em));
    return CHIML_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------
// commit helpers
// ---------------------------------------------------------------------------------------------------
namespace {

// lanes per emitter of the density kernel (chiml_emitters.cuh): 0 = one thread per emitter (N = 2, 3: no or small spills; N = 6: more
// elements than a warp has lanes), else the group kernel (N = 4, 5).  CHIML_B200_EMIT_KERNEL=thread|group overrides (tests, A/B).
int emitter_group(int nlevel)
{
    const int g = nlevel == 2 ? 4 : (nlevel <= 4 ? 16 : 32);
    if(nlevel > 5) return 0;
    if(const char* ev = std::getenv("CHIML_B200_EMIT_KERNEL")) return std::strcmp(ev, "group") == 0 ? g : 0;
    return nlevel >= 4 ? g : 0;
}

// algorithmic bytes one field-component cell with this info value moves per step (the host twin of cell_alg_bytes, chiml_update.cuh)
double info_alg_bytes(const unsigned info, const ClassEntry* cls, const bool isE, const bool pmlOnD)
{
    if(info == 0) return 0.0;
    const bool pml = (info & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
    double b = 0.0;
    if((info & F_CURL) || (info & (F_PG0 | F_PG1))) b += 24;
    if(isE && ((info & (F_ISD | F_D2E | F_ORD2E)) || (pmlOnD && pml))) b += 16;
    if(isE && (info & F_D2E)) b += 24.0 * cls[info & CLS_MASK].npoles;
    if(info & F_PS0) b += 16;
    if(info & F_PS1) b += 16;
    return b;
}

struct ClassKey
{
    double pf1, pf2, eps; int npoles, ordip; std::vector<double> consts;
    bool operator<(const ClassKey& o) const
    {
        return std::tie(pf1, pf2, eps, npoles, ordip, consts) < std::tie(o.pf1, o.pf2, o.eps, o.npoles, o.ordip, o.consts);
    }
};

struct ClassBuilder
{
    std::map<ClassKey, int> ids;
    std::vector<ClassEntry> entries;
    ClassBuilder() { ClassEntry z; std::memset(&z, 0, sizeof(z)); entries.push_back(z); }
    // returns 0 on overflow
    // magnetic: the run belongs to an H component: its poles are the object's magnetic ones (magAlpha, magXi, magGamma)
    int get(const ChimlRun& r, const HostObj& o, bool withPoles, bool magnetic = false)
    {
        ClassKey k;
        k.pf1 = r.pf[1]; k.pf2 = r.pf[2]; k.eps = r.pf[3];
        k.npoles = withPoles ? (magnetic ? o.nmag : o.npoles) : 0;
        k.ordip = (withPoles && !magnetic) ? o.use_or_dip : 0;
        if(withPoles && magnetic)
        {
            k.consts = o.malpha;
            k.consts.insert(k.consts.end(), o.mxi.begin(), o.mxi.end());
            k.consts.insert(k.consts.end(), o.mgamma.begin(), o.mgamma.end());
        }
        else if(withPoles)
        {
            k.consts = o.alpha;
            k.consts.insert(k.consts.end(), o.xi.begin(), o.xi.end());
            k.consts.insert(k.consts.end(), o.gamma.begin(), o.gamma.end());
            k.consts.insert(k.consts.end(), o.dip.begin(), o.dip.end());
        }
        // a chiral object's D / B run and its chiral run cover the same cells: both carry the chiral constants
        const int nchi = withPoles ? o.nchi : 0;
        if(nchi > 0)
        {
            k.consts.push_back(1e300);          // (separator: the lists above have variable length)
            k.consts.insert(k.consts.end(), o.calpha.begin(), o.calpha.end());
            k.consts.insert(k.consts.end(), o.cxi.begin(), o.cxi.end());
            k.consts.insert(k.consts.end(), o.cgamma.begin(), o.cgamma.end());
            k.consts.insert(k.consts.end(), o.cgprev.begin(), o.cgprev.end());
        }
        auto it = ids.find(k);
        if(it != ids.end()) return it->second;
        if((int)entries.size() > MAX_CLASSES) return 0;
        ClassEntry e; std::memset(&e, 0, sizeof(e));
        e.pf1 = k.pf1; e.pf2 = k.pf2;
        e.inv_eps = 1.0 / k.eps; e.neg_inv_eps = -1.0 / k.eps; e.neg_half_inv_eps = -0.5 / k.eps;
        e.npoles = k.npoles;
        for(int p = 0; p < k.npoles; ++p)
        {
            if(magnetic) { e.alpha[p] = o.malpha[p]; e.xi[p] = o.mxi[p]; e.gamma[p] = o.mgamma[p]; continue; }
            e.alpha[p] = o.alpha[p]; e.xi[p] = o.xi[p]; e.gamma[p] = o.gamma[p];
            for(int q = 0; q < 3; ++q) e.dip[p][q] = o.dip[3 * p + q];
        }
        e.nchi = nchi;
        for(int p = 0; p < nchi; ++p)
        {
            e.chi_alpha[p] = o.calpha[p]; e.chi_xi[p] = o.cxi[p];
            e.chi_g8[p] = o.cgamma[p] / 8.0; e.chi_gp8[p] = o.cgprev[p] / 8.0;      // daxpy_(n, gamma[pp] / 8.0, ...)
        }
        e.chi_fac = magnetic ? -1.0 / k.eps : -1.0 / (-1.0 * k.eps);               // chiDtoU: -1.0 / epMuInfty, epMuInfty = +mu (B2H) or -eps (D2E)
        int id = (int)entries.size();
        entries.push_back(e);
        ids[k] = id;
        return id;
    }
};

int paint_list(ChimlCtx* ctx, const std::vector<ChimlRun>& runs, const std::vector<uint8_t>& cls, uint16_t flags, uint16_t* d_info, int* d_err)
{
    if(runs.empty()) return 0;
    ChimlRun* d_runs = nullptr; uint8_t* d_cls = nullptr;
    CK(cudaMalloc((void**)&d_runs, runs.size() * sizeof(ChimlRun)));
    CK(cudaMalloc((void**)&d_cls, cls.size()));
    CK(cudaMemcpyAsync(d_runs, runs.data(), runs.size() * sizeof(ChimlRun), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_cls, cls.data(), cls.size(), cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    const size_t blocks = std::min<size_t>((runs.size() * 32 + threads - 1) / threads, 148 * 16);
    k_paint_runs<<<(unsigned)blocks, threads, 0, ctx->stream>>>(d_runs, d_cls, runs.size(), flags, d_info, ctx->lx, ctx->px, d_err);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_runs); cudaFree(d_cls);
    return 0;
}

int paint_lines(ChimlCtx* ctx, const std::vector<int4>& lines, uint16_t flags, uint16_t* d_info, int* d_err)
{
    if(lines.empty()) return 0;
    int4* d_lines = nullptr;
    CK(cudaMalloc((void**)&d_lines, lines.size() * sizeof(int4)));
    CK(cudaMemcpyAsync(d_lines, lines.data(), lines.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    const size_t blocks = std::min<size_t>((lines.size() * 32 + threads - 1) / threads, 148 * 16);
    k_paint_lines<<<(unsigned)blocks, threads, 0, ctx->stream>>>(d_lines, lines.size(), flags, d_info, ctx->lx, ctx->px, d_err);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_lines);
    return 0;
}

// per-row x-spans of the runs whose object carries poles
int build_spans(ChimlCtx* ctx, const std::vector<const std::vector<ChimlRun>*>& lists, bool ordipOnly, SpanTable& sp, bool magnetic = false)
{
    const size_t nrows = (size_t)ctx->ly * ctx->lz;
    sp.h_xmin.assign(nrows, -1); sp.h_xmax.assign(nrows, -1); sp.h_base.assign(nrows, 0);
    for(auto* l : lists)
        for(const ChimlRun& r : *l)
        {
            const HostObj& o = ctx->objs[r.obj];
            if((magnetic ? o.nmag : o.npoles) == 0 && o.nchi == 0) continue;
            if(ordipOnly && !o.use_or_dip) continue;
            const size_t row = (size_t)(r.ind / ctx->lx);
            const int x0 = r.ind % ctx->lx, x1 = x0 + r.n - 1;
            if(sp.h_xmin[row] < 0 || x0 < sp.h_xmin[row]) sp.h_xmin[row] = x0;
            if(x1 > sp.h_xmax[row]) sp.h_xmax[row] = x1;
        }
    // spans start at an even x and have an even length, so that the pool offset of an even x is even: the kernels access the
    // pools two cells (16 bytes, aligned) at a time.  The padding cells are never updated and stay zero.
    int64_t total = 0;
    for(size_t row = 0; row < nrows; ++row)
        if(sp.h_xmin[row] >= 0)
        {
            sp.h_xmin[row] &= ~1;
            sp.h_xmax[row] |= 1;
            sp.h_base[row] = total; total += sp.h_xmax[row] - sp.h_xmin[row] + 1;
        }
    sp.total = total;
    std::vector<int32_t> rows;
    sp.max_width = 0;
    for(size_t row = 0; row < nrows; ++row)
        if(sp.h_xmin[row] >= 0)
        {
            rows.push_back((int32_t)row);
            sp.max_width = std::max(sp.max_width, sp.h_xmax[row] - sp.h_xmin[row] + 1);
        }
    sp.nrows_used = (int)rows.size();
    int rc;
    if((rc = dev_upload(ctx, &sp.d_rows, rows))) return rc;
    if((rc = dev_upload(ctx, &sp.d_xmin, sp.h_xmin))) return rc;
    if((rc = dev_upload(ctx, &sp.d_xmax, sp.h_xmax))) return rc;
    if((rc = dev_upload(ctx, &sp.d_base, sp.h_base))) return rc;
    return 0;
}

int coord_of(const ChimlCtx* ctx, long ind, int axis)
{
    if(axis == 0) return (int)(ind % ctx->lx);
    if(axis == 2) return (int)((ind / ctx->lx) % ctx->lz);
    return (int)(ind / ((long)ctx->lx * ctx->lz));
}
int stride_axis(const ChimlCtx* ctx, int stride)
{
    if(stride == 1) return 0;
    if(ctx->lz > 1 && stride == ctx->lx) return 2;
    if(stride == ctx->lx * ctx->lz) return 1;
    return -1;
}

// Build the device form of one CPML part from the reference's two lists.
int build_pml_part(ChimlCtx* ctx, int comp, int part, int* d_err)
{
    const HostPml& h = ctx->hpml[comp][part];
    PmlPartDev& pp = ctx->pml[comp][part];
    if(!h.present || (h.psi.empty() && h.grid.empty())) return 0;
    const int i = comp % 3;
    const bool isE = comp < 3;
    // part 0 is driven by grid_k, part 1 by grid_j (PML/parallelPML.hpp:695-696); E components are driven by H and vice versa
    pp.vfield = (isE ? CHIML_HX : CHIML_EX) + (part == 0 ? (i + 2) % 3 : (i + 1) % 3);
    if(!field_exists(ctx, pp.vfield)) return fail(ctx, CHIML_ERR_ARG, "set_cpml: the field driving this CPML part does not exist in this mode");
    const int ln[3] = {ctx->lx, ctx->ly, ctx->lz};
    const long ncell = (long)ctx->nlogical;
    // stencil offset and derivative axis: identical for every entry of both lists
    bool haveOff = false;
    auto checkOff = [&](long off) -> bool {
        if(!haveOff) { pp.off_logical = off; haveOff = true; return true; }
        return pp.off_logical == off;
    };
    for(const auto& e : h.psi)  if(!checkOff((long)e.indOff - e.ind)) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: psi entries with different stencil offsets");
    for(const auto& e : h.grid) if(!checkOff((long)e.indOff - e.ind)) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: grid entries with different stencil offsets");
    int dd[3];
    if(!decode_offset(ctx, pp.off_logical, dd) || pp.off_logical == 0) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: stencil offset is not one cell along an axis");
    pp.axis = dd[0] ? 0 : (dd[1] ? 1 : 2);
    // part 0 differentiates along j = (i+1)%3, part 1 along k = (i+2)%3 (PML/parallelPML.hpp:124-141,679,687); the kernel relies on it
    if(pp.axis != (i + 1 + part) % 3) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: derivative axis of the list does not match the part (j for part 0, k for part 1)");
    const int L = ln[pp.axis];
    std::vector<double> F(L, 0.0), b(L, 0.0), c(L, 0.0);
    std::vector<char> Fset(L, 0), bset(L, 0);
    pp.has_psi = h.has_psi;
    pp.present = 1;
    bool dbSet = false;
    std::vector<int4> psiLines, gridLines;
    for(const auto& e : h.psi)
    {
        const int sa = stride_axis(ctx, e.stride);
        if(e.transSz < 1 || sa < 0 || e.ind < 0 || (long)e.ind + (long)(e.transSz - 1) * e.stride >= ncell)
            return fail(ctx, CHIML_ERR_ARG, "set_cpml: psi entry outside the grid");
        if(sa == pp.axis) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: psi line runs along its own normal axis");
        const int co = coord_of(ctx, e.ind, pp.axis);
        if(bset[co] && (b[co] != e.b || c[co] != e.c)) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: psi coefficients are not a function of the coordinate along the slab normal");
        b[co] = e.b; c[co] = e.c; bset[co] = 1;
        psiLines.push_back(make_int4(e.transSz, e.stride, e.ind, 0));
    }
    for(const auto& e : h.grid)
    {
        const int sa = stride_axis(ctx, e.stride);
        if(e.nAx < 1 || sa < 0 || e.ind < 0 || (long)e.ind + (long)(e.nAx - 1) * e.stride >= ncell)
            return fail(ctx, CHIML_ERR_ARG, "set_cpml: grid entry outside the grid");
        if(dbSet && pp.Db != e.Db) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: Db differs between grid entries");
        pp.Db = e.Db; dbSet = true;
        const int co0 = coord_of(ctx, e.ind, pp.axis);
        const int nco = sa == pp.axis ? e.nAx : 1;
        for(int k = 0; k < nco; ++k)
        {
            const int co = co0 + k;
            if(co >= L) return fail(ctx, CHIML_ERR_ARG, "set_cpml: grid entry leaves the grid");
            if(Fset[co] && F[co] != e.DbField) return fail(ctx, CHIML_ERR_UNSUPPORTED, "set_cpml: DbField is not a function of the coordinate along the derivative axis");
            F[co] = e.DbField; Fset[co] = 1;
        }
        gridLines.push_back(make_int4(e.nAx, e.stride, e.ind, 0));
    }
    // compact psi coordinates
    pp.h_cmap.assign(L, -1);
    pp.nact = 0;
    for(int co = 0; co < L; ++co) if(bset[co]) pp.h_cmap[co] = pp.nact++;
    int rc;
    if((rc = dev_upload(ctx, &pp.d_F, F))) return rc;
    if((rc = dev_upload(ctx, &pp.d_b, b))) return rc;
    if((rc = dev_upload(ctx, &pp.d_c, c))) return rc;
    if((rc = dev_upload(ctx, &pp.d_cmap, pp.h_cmap))) return rc;
    if(pp.has_psi && pp.nact > 0)
    {
        if(pp.axis == 0) { pp.psi_pitch = ((long)pp.nact + 1) / 2 * 2; pp.psi_count = pp.psi_pitch * ctx->lz * (long)ctx->ly; }
        else if(pp.axis == 1) { pp.psi_pitch = ctx->px; pp.psi_count = ctx->px * ctx->lz * (long)pp.nact; }
        else { pp.psi_pitch = ctx->px; pp.psi_count = ctx->px * (long)pp.nact * ctx->ly; }
        if((rc = dev_alloc(ctx, &pp.d_psi, (size_t)pp.psi_count))) return rc;
    }
    if((rc = paint_lines(ctx, psiLines, part == 0 ? F_PS0 : F_PS1, ctx->d_info[comp], d_err))) return rc;
    if((rc = paint_lines(ctx, gridLines, part == 0 ? F_PG0 : F_PG1, ctx->d_info[comp], d_err))) return rc;
    return 0;
}

// TMA descriptors of one-row boxes (64 x 1 x 1 doubles) of every field array and psi array, for the row prefetch of the UNIFORM
// marching kernels (chiml_update.cuh tma_prefetch_row).  cuTensorMapEncodeTiled is looked up in the driver at run time; when it is
// missing (or CHIML_B200_NO_TMA is set) the kernels fall back to per-lane prefetch.global.L2.
int build_tensor_maps(ChimlCtx* ctx)
{
    if(ctx->lz <= 1 || std::getenv("CHIML_B200_NO_TMA")) return 0;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess)
    { cudaGetLastError(); return 0; }
    EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
    static_assert(sizeof(CUtensorMap) == TMAP_BYTES, "CUtensorMap is 128 bytes");
    std::vector<CUtensorMap> maps(TMAP_COUNT);
    std::memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
    auto make = [&](CUtensorMap& m, double* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t pitch0, cuuint64_t pitch1) -> bool {
        const cuuint64_t dims[3] = {d0, d1, d2};
        const cuuint64_t strides[2] = {pitch0 * sizeof(double), pitch1 * sizeof(double)};       // of dimensions 1 and 2, in bytes
        const cuuint32_t box[3] = {(cuuint32_t)std::min<cuuint64_t>(d0, TILE_X), 1, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        return encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = true;
    for(int f = 0; f < TMAP_PSI0 && ok; ++f)
        if(ctx->d_field[f]) ok = make(maps[f], ctx->d_field[f], (cuuint64_t)ctx->px, (cuuint64_t)ctx->lz, (cuuint64_t)ctx->ly, (cuuint64_t)ctx->px, (cuuint64_t)ctx->plane);
    for(int comp = 0; comp < 6 && ok; ++comp)
        for(int part = 0; part < 2 && ok; ++part)
        {
            const PmlPartDev& pp = ctx->pml[comp][part];
            if(!pp.d_psi) continue;
            CUtensorMap& m = maps[TMAP_PSI0 + 2 * comp + part];
            // the compact layouts of build_pml_part: x-normal slabs cc + pitch * (z + lz * y); y-normal x + px * (z + lz * cc); z-normal x + px * (cc + nact * y)
            if(pp.axis == 0)      ok = make(m, pp.d_psi, (cuuint64_t)pp.psi_pitch, (cuuint64_t)ctx->lz * ctx->ly, 1, (cuuint64_t)pp.psi_pitch, (cuuint64_t)pp.psi_pitch * ctx->lz * ctx->ly);
            else if(pp.axis == 1) ok = make(m, pp.d_psi, (cuuint64_t)ctx->px, (cuuint64_t)ctx->lz, (cuuint64_t)pp.nact, (cuuint64_t)ctx->px, (cuuint64_t)ctx->plane);
            else                  ok = make(m, pp.d_psi, (cuuint64_t)ctx->px, (cuuint64_t)pp.nact, (cuuint64_t)ctx->ly, (cuuint64_t)ctx->px, (cuuint64_t)ctx->px * pp.nact);
        }
    if(!ok) return 0;            // a layout the encoder refuses (e.g. a stride that is no multiple of 16 bytes): per-lane prefetch instead
    CK(cudaMalloc(&ctx->d_tmaps, maps.size() * sizeof(CUtensorMap)));
    ctx->dev_bytes += maps.size() * sizeof(CUtensorMap);
    CK(cudaMemcpyAsync(ctx->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

} // namespace

extern "C" {

int chiml_gpu_commit(ChimlCtx* ctx)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(ctx->committed) return fail(ctx, CHIML_ERR_STATE, "commit called twice");
    CK(cudaSetDevice(ctx->device));
    if(ctx->periodic)
    {
        // the reference also wraps the node-centred oriented-dipole pole grids (applBCOrDip_, parallelFDTDField.hpp:1362-1363): the compact
        // node pools hold no ghost nodes
        bool ordip = !ctx->lists[CHIML_LIST_ORDIPP][0].runs.empty();
        for(int c = 0; c < 3; ++c) ordip = ordip || !ctx->lists[CHIML_LIST_ORDIPD][c].runs.empty();
        if(ordip) return fail(ctx, CHIML_ERR_UNSUPPORTED, "oriented-dipole media under periodic boundaries are outside the covered hot path");
        for(int comp = 0; comp < 6; ++comp)
            if(field_exists(ctx, comp) && !ctx->has_wrap[comp]) return fail(ctx, CHIML_ERR_ARG, "periodic boundaries: set_periodic must be called for every field component");
        // a ring of slabs carries the field rows the curls read; emitter polarisations, B / M and TFSF corrections across the seam are not built
        if(ctx->ring && (ctx->has_B || !ctx->tfsf.empty()))
            return fail(ctx, CHIML_ERR_UNSUPPORTED, "periodic runs on several slabs are covered without magnetic / chiral media and TFSF surfaces");
        if(ctx->ring && ctx->ly < 5) return fail(ctx, CHIML_ERR_UNSUPPORTED, "periodic runs on several slabs need at least three owned rows per slab");
    }
    int rc;
    int* d_err = nullptr;
    if((rc = dev_alloc(ctx, &d_err, 1))) return rc;

    // fields (+1 padded row of slack on either side is not needed: every stencil point of an updated cell lies inside the ghost-inclusive grid)
    // one plane (+ a row) of zeroed slack on either side: the kernels load the stencil neighbours of ghost and
    // padding cells unconditionally
    ctx->guard = (size_t)ctx->plane + 32;
    for(int f = 0; f < CHIML_NFIELDS; ++f)
        if(field_exists(ctx, f))
        {
            if((rc = dev_alloc_ipc(ctx, &ctx->d_field_base[f], ctx->nphys + 2 * ctx->guard))) return rc;
            ctx->d_field[f] = ctx->d_field_base[f] + ctx->guard;
        }

    // chiral media: prevE_ / prevH_ and the rows copied into them
    {
        bool anyChi = false;
        for(int comp = 0; comp < 6; ++comp) anyChi = anyChi || !ctx->lists[CHIML_LIST_CHID][comp].runs.empty();
        if(anyChi)
        {
            for(int c = 0; c < 6; ++c)
            {
                if((rc = dev_alloc(ctx, &ctx->d_prev_base[c], ctx->nphys + 2 * ctx->guard))) return rc;
                ctx->d_prev[c] = ctx->d_prev_base[c] + ctx->guard;
            }
            std::vector<int4> rows(ctx->h_prev_rows.size() / 4);
            for(size_t q = 0; q < rows.size(); ++q) rows[q] = make_int4(ctx->h_prev_rows[4 * q], ctx->h_prev_rows[4 * q + 1], ctx->h_prev_rows[4 * q + 2], ctx->h_prev_rows[4 * q + 3]);
            ctx->n_prev_rows = rows.size();
            if((rc = dev_upload(ctx, &ctx->d_prev_rows, rows))) return rc;
        }
    }

    // --- update lists -> class tables + painted cell info -------------------------------------------------
    for(int comp = 0; comp < 6; ++comp)
    {
        if(!field_exists(ctx, comp))
        {
            for(int k = 0; k < 6; ++k)
                if(k != CHIML_LIST_ORDIPP && !ctx->lists[k][comp].runs.empty())
                    return fail(ctx, CHIML_ERR_ARG, "update list given for a field component that does not exist in this mode");
            continue;
        }
        if((rc = dev_alloc(ctx, &ctx->d_info[comp], ctx->nphys))) return rc;
        ClassBuilder cb;
        const bool isE = comp < 3;
        if(isE ? !ctx->g.has_D : !ctx->has_B)
            for(int k : {CHIML_LIST_D, CHIML_LIST_LORD, CHIML_LIST_ORDIPD, CHIML_LIST_CHID})
                if(!ctx->lists[k][comp].runs.empty()) return fail(ctx, CHIML_ERR_ARG, isE ? "D-type update list given but has_D = 0" : "B-type update list given but has_B = 0");
        // stencil offsets must be uniform per component
        bool& haveOff = ctx->have_off[comp];
        struct { int kind; uint16_t flags; bool poles; } plan[5] = {
            {CHIML_LIST_U, F_CURL, false}, {CHIML_LIST_D, (uint16_t)(F_CURL | F_ISD), true},
            {CHIML_LIST_LORD, F_D2E, true}, {CHIML_LIST_ORDIPD, F_ORD2E, true}, {CHIML_LIST_CHID, F_D2E, true}};
        for(auto& pl : plan)
        {
            const auto& runs = ctx->lists[pl.kind][comp].runs;
            if(runs.empty()) continue;
            std::vector<uint8_t> cls(runs.size());
            for(size_t e = 0; e < runs.size(); ++e)
            {
                const ChimlRun& r = runs[e];
                if(pl.kind == CHIML_LIST_U || pl.kind == CHIML_LIST_D)
                {
                    const long oj = (long)r.ind_j - r.ind, ok = (long)r.ind_k - r.ind;
                    if(!haveOff) { ctx->off_j[comp] = oj; ctx->off_k[comp] = ok; haveOff = true; }
                    else if(ctx->off_j[comp] != oj || ctx->off_k[comp] != ok)
                        return fail(ctx, CHIML_ERR_UNSUPPORTED, "update list: stencil offsets differ between runs of one component");
                }
                if(pl.kind == CHIML_LIST_CHID)
                {
                    // the eight-point stencil of UpdateChiral hangs on the entry's three neighbour indices: one geometry per component
                    const long o3[3] = {(long)r.ind_i - r.ind, (long)r.ind_j - r.ind, (long)r.ind_k - r.ind};
                    for(int q = 0; q < 3; ++q)
                    {
                        if(e == 0) ctx->chi_off[comp][q] = o3[q];
                        else if(ctx->chi_off[comp][q] != o3[q]) return fail(ctx, CHIML_ERR_UNSUPPORTED, "chiral list: neighbour offsets differ between runs");
                    }
                    if(ctx->objs[r.obj].nchi == 0) return fail(ctx, CHIML_ERR_ARG, "chiral list: the object of a run has no chiral pole (chiml_gpu_set_object_chiral)");
                    ctx->has_chi = true;
                }
                if(pl.kind == CHIML_LIST_ORDIPD)
                {
                    const long oi = (long)r.ind_i - r.ind;
                    if(e == 0) ctx->ordip_off[comp] = oi;
                    else if(ctx->ordip_off[comp] != oi) return fail(ctx, CHIML_ERR_UNSUPPORTED, "oriented-dipole list: node offsets differ between runs");
                }
                if(r.pf[3] == 0.0 && pl.kind != CHIML_LIST_U) return fail(ctx, CHIML_ERR_ARG, "update list: eps = 0 in a D-type run");
                const HostObj& o = ctx->objs[r.obj];
                // E-side D-type runs of isotropic objects carry the object's poles in their class (so that the
                // D-run and the LorD-run of one cell agree); oriented-dipole objects keep theirs on the node grid
                const bool usePoles = isE ? (pl.poles && !o.use_or_dip) : (pl.poles && ctx->has_B);
                int id = cb.get(r, o, usePoles, !isE);
                if(id == 0) return fail(ctx, CHIML_ERR_UNSUPPORTED, "more than 255 distinct material classes for one field component");
                cls[e] = (uint8_t)id;
            }
            if((rc = paint_list(ctx, runs, cls, pl.flags, ctx->d_info[comp], d_err))) return rc;
        }
        ctx->ncls[comp] = (int)cb.entries.size();
        ctx->h_cls[comp] = cb.entries;
        if((rc = dev_upload(ctx, &ctx->d_cls[comp], cb.entries))) return rc;
        std::vector<double2> pf(cb.entries.size());
        for(size_t k = 0; k < pf.size(); ++k) pf[k] = make_double2(cb.entries[k].pf1, cb.entries[k].pf2);
        if((rc = dev_upload(ctx, &ctx->d_pf[comp], pf))) return rc;
        int dj[3], dk[3];
        if(haveOff && (!decode_offset(ctx, ctx->off_j[comp], dj) || !decode_offset(ctx, ctx->off_k[comp], dk)))
            return fail(ctx, CHIML_ERR_UNSUPPORTED, "update list: stencil offset is not one cell along an axis");
    }
    // a D-run and a LorD/OrDipD run covering the same cell must agree on the class byte: the D-run was keyed with poles too
    // --- isotropic pole pools -----------------------------------------------------------------------------
    for(int c = 0; c < 6; ++c)
    {
        if(!field_exists(ctx, c) || (c < 3 ? !ctx->g.has_D : !ctx->has_B)) continue;
        int np = 0, nchi = 0;
        for(int kind : {CHIML_LIST_LORD, CHIML_LIST_CHID})
            for(const ChimlRun& r : ctx->lists[kind][c].runs)
            {
                if(c >= 3) np = std::max(np, ctx->objs[r.obj].nmag);
                else if(!ctx->objs[r.obj].use_or_dip) np = std::max(np, ctx->objs[r.obj].npoles);
                if(kind == CHIML_LIST_CHID) nchi = std::max(nchi, ctx->objs[r.obj].nchi);
            }
        ctx->npoles_comp[c] = np;
        ctx->nchi_comp[c] = nchi;
        std::vector<const std::vector<ChimlRun>*> ls = {&ctx->lists[CHIML_LIST_LORD][c].runs, &ctx->lists[CHIML_LIST_CHID][c].runs};
        if((rc = build_spans(ctx, ls, false, ctx->span[c], c >= 3))) return rc;
        for(int p = 0; p < np; ++p)
            for(int k = 0; k < 2; ++k)
                if((rc = dev_alloc(ctx, &ctx->d_P[c][p][k], (size_t)ctx->span[c].total))) return rc;
        for(int p = 0; p < nchi; ++p)
            for(int k = 0; k < 2; ++k)
                if((rc = dev_alloc(ctx, &ctx->d_chi[c][p][k], (size_t)ctx->span[c].total))) return rc;
    }
    // --- oriented-dipole node grid ------------------------------------------------------------------------
    {
        const auto& runs = ctx->lists[CHIML_LIST_ORDIPP][0].runs;
        ctx->nordip = ctx->g.nranks > 1 ? ctx->nordip_global : 0;
        for(const ChimlRun& r : runs) ctx->nordip = std::max(ctx->nordip, ctx->objs[r.obj].npoles);
        if(!runs.empty())
        {
            if(!ctx->g.has_D) return fail(ctx, CHIML_ERR_ARG, "oriented-dipole node list given but has_D = 0");
            if((rc = dev_alloc(ctx, &ctx->d_info_node, ctx->nphys))) return rc;
            ClassBuilder cb;
            std::vector<uint8_t> cls(runs.size());
            for(size_t e = 0; e < runs.size(); ++e)
            {
                const ChimlRun& r = runs[e];
                const long o3[3] = {(long)r.ind_i - r.ind, (long)r.ind_j - r.ind, (long)r.ind_k - r.ind};
                for(int k = 0; k < 3; ++k)
                {
                    if(e == 0) ctx->node_off[k] = o3[k];
                    else if(ctx->node_off[k] != o3[k]) return fail(ctx, CHIML_ERR_UNSUPPORTED, "oriented-dipole node list: offsets differ between runs");
                }
                const HostObj& o = ctx->objs[r.obj];
                ChimlRun key = r; key.pf[1] = key.pf[2] = 0.0; key.pf[3] = 1.0;   // node classes depend on the pole constants only
                int id = cb.get(key, o, true);
                if(id == 0) return fail(ctx, CHIML_ERR_UNSUPPORTED, "more than 255 distinct oriented-dipole material classes");
                cls[e] = (uint8_t)id;
                int ncomp = 0;
                for(int c = 0; c < 3; ++c) ncomp += field_exists(ctx, c) ? 1 : 0;
                ctx->kstat[K_ORDIP_POLES].alg_bytes += 24.0 * ncomp * (double)r.n * o.npoles;   // P, prevP read, new P written
            }
            if((rc = paint_list(ctx, runs, cls, 0x0100, ctx->d_info_node, d_err))) return rc;
            ctx->ncls_node = (int)cb.entries.size();
            if((rc = dev_upload(ctx, &ctx->d_cls_node, cb.entries))) return rc;
            std::vector<const std::vector<ChimlRun>*> ls = {&runs};
            if((rc = build_spans(ctx, ls, true, ctx->span_node))) return rc;
            for(int c = 0; c < 3; ++c)
                if(field_exists(ctx, c))
                    for(int p = 0; p < ctx->nordip; ++p)
                        for(int k = 0; k < 2; ++k)
                            if((rc = dev_alloc(ctx, &ctx->d_oP[c][p][k], (size_t)ctx->span_node.total))) return rc;
            // position-dependent dipole grids: the values at the node cells, in pool order (cells of a span the list does not cover are never read)
            for(int c = 0; c < 3; ++c)
                for(int p = 0; p < MAX_POLES; ++p)
                {
                    const double* hg = ctx->h_dipg[c][p];
                    if(!hg) continue;
                    if(p < ctx->nordip && field_exists(ctx, c))
                    {
                        const SpanTable& sp = ctx->span_node;
                        std::vector<double> tmp((size_t)sp.total, 0.0);
                        const size_t nrows = (size_t)ctx->ly * ctx->lz;
                        for(size_t row = 0; row < nrows; ++row)
                            if(sp.h_xmin[row] >= 0)
                                std::copy_n(hg + row * ctx->lx + sp.h_xmin[row], std::min(sp.h_xmax[row], ctx->lx - 1) - sp.h_xmin[row] + 1, &tmp[(size_t)sp.h_base[row]]);
                        if((rc = dev_upload(ctx, &ctx->d_dipg[c][p], tmp))) return rc;
                        // (static data: not part of the algorithmic bytes, like the class tables and CPML coefficients)
                    }
                    ctx->h_dipg[c][p] = nullptr;
                }
        }
        else if(ctx->nordip == 0)
            // (with several slabs a slab may hold edge cells of an object whose nodes all lie in the slab above: its D->E then reads
            // the ghost row only, and the global pole count says how many)
            for(int c = 0; c < 3; ++c)
                if(!ctx->lists[CHIML_LIST_ORDIPD][c].runs.empty())
                    return fail(ctx, CHIML_ERR_ARG, "oriented-dipole D->E list given without the node list");
    }
    // --- CPML ---------------------------------------------------------------------------------------------
    for(int comp = 0; comp < 6; ++comp)
        for(int part = 0; part < 2; ++part)
        {
            if(ctx->hpml[comp][part].present && !field_exists(ctx, comp))
                return fail(ctx, CHIML_ERR_ARG, "CPML lists given for a field component that does not exist in this mode");
            if(field_exists(ctx, comp) && (rc = build_pml_part(ctx, comp, part, d_err))) return rc;
        }
    if(ctx->g.pml_on_D && !ctx->g.has_D) return fail(ctx, CHIML_ERR_ARG, "pml_on_D needs has_D");
    // The kernels feed the CPML parts with the stencil values of the curl: part 0 (grid_k, derivative along j) must use
    // the offset of ind_j, part 1 (grid_j, derivative along k) the offset of ind_k (true for the reference: both follow derivOff)
    for(int comp = 0; comp < 6; ++comp)
    {
        long* offs[2] = {&ctx->off_j[comp], &ctx->off_k[comp]};
        for(int part = 0; part < 2; ++part)
        {
            const PmlPartDev& pp = ctx->pml[comp][part];
            if(!pp.present) continue;
            if(!ctx->have_off[comp] && *offs[part] == 0) *offs[part] = pp.off_logical;
            else if(*offs[part] != pp.off_logical)
                return fail(ctx, CHIML_ERR_UNSUPPORTED, "CPML stencil offset differs from the curl stencil offset of the same component");
        }
    }

    int h_err = 0;
    CK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_err);
    if(h_err == 1) return fail(ctx, CHIML_ERR_UNSUPPORTED, "lists disagree on the material of a cell (different eps / prefactors for the same cell)");
    if(h_err == 2) return fail(ctx, CHIML_ERR_ARG, "an update list covers a cell twice");
    if(h_err == 3) return fail(ctx, CHIML_ERR_ARG, "a CPML list covers a cell twice");

    // The kernels hard-wire the reference's stencil (derivOff, FDTD_MANAGER/parallelFDTDField.cpp:80-82,248-250):
    // component c reads grid_k one cell along axis j=(c+1)%3 (ind_j) and grid_j one cell along axis k=(c+2)%3 (ind_k),
    // backwards for E and forwards for H.  Refuse lists that say otherwise.
    {
        const long strideL[3] = {1, (long)ctx->lx * ctx->lz, ctx->lx};   // logical strides of x, y, z
        for(int comp = 0; comp < 6; ++comp)
        {
            if(!field_exists(ctx, comp)) continue;
            const int i = comp % 3, s = comp < 3 ? -1 : 1, base = comp < 3 ? CHIML_HX : CHIML_EX;
            const bool hasVj = field_exists(ctx, base + (i + 1) % 3), hasVk = field_exists(ctx, base + (i + 2) % 3);
            const bool anyList = ctx->have_off[comp] || ctx->pml[comp][0].present || ctx->pml[comp][1].present;
            if(!anyList) continue;
            if(hasVk && ctx->off_j[comp] != 0 && ctx->off_j[comp] != s * strideL[(i + 1) % 3])
                return fail(ctx, CHIML_ERR_UNSUPPORTED, "update list: ind_j is not the reference's Yee stencil neighbour");
            if(hasVj && ctx->off_k[comp] != 0 && ctx->off_k[comp] != s * strideL[(i + 2) % 3])
                return fail(ctx, CHIML_ERR_UNSUPPORTED, "update list: ind_k is not the reference's Yee stencil neighbour");
        }
    }
    // tile classification -> compact work lists for k_fast / k_uniform / k_general (chiml_update.cuh)
    {
        const dim3 tb(32, ctx->lz > 1 ? TILE_Z : 1, 1);
        const unsigned nxt = (ctx->lx + TILE_X - 1) / TILE_X, nzt = (ctx->lz + tb.y - 1) / tb.y;
        const size_t ntiles = (size_t)nxt * nzt * ctx->ly;
        TileSummary* d_sum = nullptr;
        CK(cudaMalloc((void**)&d_sum, ntiles * sizeof(TileSummary)));
        std::vector<TileSummary> sum(ntiles);
        for(int fam = 0; fam < 2; ++fam)
        {
            const int b0 = fam == 0 ? 0 : 3;
            k_tile_summary<<<(unsigned)ntiles, tb, 0, ctx->stream>>>(ctx->d_info[b0], ctx->d_info[b0 + 1], ctx->d_info[b0 + 2], d_sum, nxt, nzt, ctx->lz, ctx->px,
                                                                     ctx->d_cls[b0], ctx->d_cls[b0 + 1], ctx->d_cls[b0 + 2], fam == 0 ? 1 : 0, ctx->g.pml_on_D);
            ++ctx->launches;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(sum.data(), d_sum, ntiles * sizeof(TileSummary), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            std::vector<TileRec> lists[3];
            double listBytes[3] = {0.0, 0.0, 0.0};
            // CHIML_B200_DEBUG_TILES=1: why tiles ended up in the GENERAL list (development aid, printed once per half step at commit)
            const bool dbgTiles = std::getenv("CHIML_B200_DEBUG_TILES") != nullptr;
            std::map<std::string, std::pair<size_t, double>> dbgWhy;
            for(size_t tIdx = 0; tIdx < ntiles; ++tIdx)
            {
                const TileSummary& ts = sum[tIdx];
                if(!(ts.total[0] | ts.total[1] | ts.total[2])) continue;
                TileRec rec;
                std::memset(&rec, 0, sizeof(rec));
                rec.x0 = (int)(tIdx % nxt) * TILE_X;
                rec.z0 = (int)((tIdx / nxt) % nzt) * (int)tb.y;
                rec.y = (int)(tIdx / ((size_t)nxt * nzt));
                rec.ny = 1;
                bool uniform = true;
                // a component's cells must fall into a few rectangles of one info value each; a record carries two of them per component,
                // so a tile cut by several CPML / material boundaries becomes several records (k_uniform blocks)
                std::vector<TileVal> vals[3];
                for(int c = 0; c < 3; ++c)
                {
                    if(!ts.total[c]) continue;
                    if(!tile_rectangles(ts, c, vals[c]))          // chiml_tiles.hpp
                    {
                        if(dbgTiles && uniform)
                        {
                            char key[200];
                            std::snprintf(key, sizeof(key), "c%d %s xt=%d zt=%d %04x(%u) %04x(%u) %04x(%u) total=%u", c, ts.other[c] ? ">6 values" : "non-rect",
                                          (int)(tIdx % nxt), (int)((tIdx / nxt) % nzt), ts.info[c][0], ts.count[c][0], ts.info[c][1], ts.count[c][1],
                                          ts.info[c][2], ts.count[c][2], ts.total[c]);
                            auto& e = dbgWhy[key]; ++e.first; e.second += ts.bytes;
                        }
                        uniform = false; continue;
                    }
                    // magnetic-dispersive cells (B targets, magnetic poles, B -> H) are worked on by the per-cell kernel only
                    if(fam == 1)
                        for(const TileVal& tv : vals[c]) if(tv.info & (F_ISD | F_D2E)) uniform = false;
                    // ... and so are the cells of chiral objects, in either family
                    if(ctx->has_chi)
                        for(const TileVal& tv : vals[c]) if(ctx->h_cls[b0 + c][tv.info & CLS_MASK].nchi > 0) uniform = false;
                }
                if(!uniform) { lists[2].push_back(rec); listBytes[2] += ts.bytes; continue; }
                const int per = rectangles_per_record(vals);
                int nrec = 1;
                for(int c = 0; c < 3; ++c) nrec = std::max(nrec, ((int)vals[c].size() + per - 1) / per);
                double tileBytes = 0.0;
                for(int k = 0; k < nrec; ++k)
                {
                    TileRec part = rec;
                    part.part = (unsigned)k;
                    // a record whose rectangles are all plain curl cells is a k_fast record (all three components per thread)
                    bool recFast = nrec == 1;
                    double recBytes = 0.0;
                    for(int c = 0; c < 3; ++c)
                        for(int w = per * k; w < std::min((int)vals[c].size(), per * k + per); ++w)
                        {
                            const unsigned inf = vals[c][w].info;
                            const ClassEntry& ce = ctx->h_cls[b0 + c][inf & CLS_MASK];
                            const unsigned npc = (fam == 0 && (inf & F_D2E)) ? (unsigned)ce.npoles << (8 * c) : 0u;    // isotropic poles of the class
                            if(w == per * k) { part.rect[c] = vals[c][w].rect; part.info[c] = inf; part.pf[c] = make_double2(ce.pf1, ce.pf2); part.inv_eps[c] = ce.inv_eps; part.np |= npc; }
                            else           { part.rectB[c] = vals[c][w].rect; part.infoB[c] = inf; part.pfB[c] = make_double2(ce.pf1, ce.pf2); part.inv_epsB[c] = ce.inv_eps; part.npB |= npc; recFast = false; }
                            if((inf & 0xFF00u) != F_CURL) recFast = false;
                            recBytes += (double)rect_area(vals[c][w].rect) * info_alg_bytes(inf, ctx->h_cls[b0 + c].data(), fam == 0, ctx->g.pml_on_D != 0);
                        }
                    lists[recFast ? 0 : 1].push_back(part);
                    listBytes[recFast ? 0 : 1] += recBytes;
                    tileBytes += recBytes;
                }
                if(tileBytes != (double)ts.bytes) return fail(ctx, CHIML_ERR_STATE, "internal: the records of a tile do not account for its cells");
            }
            // (Cutting x-runs of equal records afresh from the start of the run, so that only the last record of a run is partial, was
            // measured and dropped: record origins that are not multiples of 64 cells cost k_fast 8 % on EVERY interior tile --
            // profiles/README.md r2_07.)
            // narrow UNIFORM records of two z-adjacent tiles become one record worked on by half-warps (k_uniform, REC_WIDE2)
            if(ctx->lz > 1 && uniform_split<true>() == 1 && uniform_split<false>() == 1 && !std::getenv("CHIML_B200_NO_WIDE2"))
            {
                std::vector<TileRec>& ul = lists[1];
                std::map<std::tuple<int, int, int, unsigned>, size_t> at;
                for(size_t i = 0; i < ul.size(); ++i) at[std::make_tuple(ul[i].y, ul[i].z0, ul[i].x0, ul[i].part)] = i;
                std::vector<char> dead(ul.size(), 0);
                for(size_t i = 0; i < ul.size(); ++i)
                {
                    TileRec& p = ul[i];
                    if(dead[i] || p.z0 % (2 * TILE_Z) != 0) continue;
                    auto it = at.find(std::make_tuple(p.y, p.z0 + TILE_Z, p.x0, p.part));
                    if(it == at.end() || dead[it->second]) continue;
                    const TileRec& q = ul[it->second];
                    unsigned xlo = 255, xhi = 0;
                    bool ok = true;
                    for(int c = 0; c < 3 && ok; ++c)
                    {
                        if(p.rectB[c] || q.rectB[c] || (p.rect[c] == 0) != (q.rect[c] == 0)) { ok = false; break; }
                        if(!p.rect[c]) continue;
                        ok = (p.rect[c] & 0xFFFFu) == (q.rect[c] & 0xFFFFu) && (p.rect[c] >> 24) == (unsigned)TILE_Z && ((q.rect[c] >> 16) & 0xFFu) == 0u &&
                             p.info[c] == q.info[c] && std::memcmp(&p.pf[c], &q.pf[c], sizeof(double2)) == 0 && p.inv_eps[c] == q.inv_eps[c];
                        xlo = std::min(xlo, p.rect[c] & 0xFFu); xhi = std::max(xhi, (p.rect[c] >> 8) & 0xFFu);
                    }
                    if(!ok || p.np != q.np || xhi <= xlo || xhi - (xlo & ~1u) > 32u) continue;
                    for(int c = 0; c < 3; ++c)
                        if(p.rect[c]) p.rect[c] = (p.rect[c] & 0x00FFFFFFu) | (((unsigned)TILE_Z + (q.rect[c] >> 24)) << 24);
                    p.part |= REC_WIDE2;
                    p.pad4 = xlo & ~1u;
                    dead[it->second] = 1;
                }
                std::vector<TileRec> kept;
                kept.reserve(ul.size());
                for(size_t i = 0; i < ul.size(); ++i) if(!dead[i]) kept.push_back(ul[i]);
                ul.swap(kept);
            }
            if(dbgTiles)
            {
                std::fprintf(stderr, "[chiml tiles] %s: fast %zu (%.3f GB) uniform %zu (%.3f GB) general %zu (%.3f GB)\n", fam == 0 ? "E" : "H",
                             lists[0].size(), listBytes[0] / 1e9, lists[1].size(), listBytes[1] / 1e9, lists[2].size(), listBytes[2] / 1e9);
                std::vector<std::pair<double, std::string>> top;
                for(auto& kv : dbgWhy) top.push_back({kv.second.second, kv.first + " tiles=" + std::to_string(kv.second.first)});
                std::sort(top.rbegin(), top.rend());
                for(size_t i = 0; i < top.size() && i < 24; ++i) std::fprintf(stderr, "[chiml tiles]   %.4f GB  %s\n", top[i].first / 1e9, top[i].second.c_str());
            }
            // FAST tiles that are stacked along y with identical rectangles and prefactors become one work item: the block marches
            // over up to MARCH_NY planes carrying the y-neighbour planes in registers (k_fast).  Slab-boundary planes stay single.
            {
                // column length: long enough to amortise the carried planes, short enough that a small grid still yields several
                // work items per SM (a 512 x 512 grid has only ~4600 tiles per half step)
                // (2-D grids pack ROWS_2D one-row tiles into a block, so they need that many more columns for the same number of blocks)
                const size_t marchCap = ntiles / (148 * 8 * (ctx->lz > 1 ? 1 : ROWS_2D));
                const bool hasLo = ctx->ring || ctx->g.rank > 0, hasUp = ctx->ring || ctx->g.rank < ctx->g.nranks - 1;
                auto isBnd = [&](const TileRec& t) { return (hasLo && t.y == 1) || (hasUp && t.y == ctx->ly - 2); };
                for(int kind = 0; kind < 2; ++kind)      // FAST and UNIFORM lists
                {
                    // measured on the C5 slab (profiles/README.md r1z): 64 planes for the vacuum kernel, 32 for the UNIFORM ones
                    int MARCH_NY = (int)std::max<size_t>(1, std::min<size_t>(kind == 0 ? 64 : 32, marchCap));
                    // forced column length (chiml_gpu_set_march, else CHIML_B200_MARCH_NY="fast[,uniform]"): the parity tests use it to
                    // run the carried-plane code of the marching kernels on grids that are too small to get columns on their own
                    int forced = kind == 0 ? ctx->march_fast : ctx->march_uniform;
                    if(forced <= 0)
                        if(const char* ev = std::getenv("CHIML_B200_MARCH_NY"))
                        {
                            int f0 = 0, f1 = 0;
                            const int got = std::sscanf(ev, "%d,%d", &f0, &f1);
                            if(got >= 1) forced = (kind == 0 || got < 2) ? f0 : f1;
                        }
                    if(forced > 0) MARCH_NY = std::min(forced, 1 << 20);
                    std::vector<TileRec>& fl = lists[kind];
                    std::stable_sort(fl.begin(), fl.end(), [](const TileRec& p, const TileRec& q) {
                        return std::tie(p.z0, p.x0, p.part, p.y) < std::tie(q.z0, q.x0, q.part, q.y); });
                    std::vector<TileRec> merged;
                    for(const TileRec& t : fl)
                    {
                        if(!merged.empty())
                        {
                            TileRec& m = merged.back();
                            const bool same = m.x0 == t.x0 && m.z0 == t.z0 && m.part == t.part && m.y + m.ny == t.y && m.ny < MARCH_NY && !isBnd(m) && !isBnd(t) &&
                                              std::memcmp(m.rect, t.rect, sizeof(m.rect)) == 0 && std::memcmp(m.pf, t.pf, sizeof(m.pf)) == 0 &&
                                              std::memcmp(m.info, t.info, sizeof(m.info)) == 0 && std::memcmp(m.inv_eps, t.inv_eps, sizeof(m.inv_eps)) == 0 &&
                                              std::memcmp(m.rectB, t.rectB, sizeof(m.rectB)) == 0 && std::memcmp(m.infoB, t.infoB, sizeof(m.infoB)) == 0 &&
                                              std::memcmp(m.pfB, t.pfB, sizeof(m.pfB)) == 0 && std::memcmp(m.inv_epsB, t.inv_epsB, sizeof(m.inv_epsB)) == 0;
                            if(same) { ++m.ny; continue; }
                        }
                        merged.push_back(t);
                    }
                    // launch order: by first plane, then z, then x -- neighbouring blocks stream neighbouring rows
                    std::stable_sort(merged.begin(), merged.end(), [](const TileRec& p, const TileRec& q) {
                        return std::tie(p.y, p.z0, p.x0) < std::tie(q.y, q.z0, q.x0); });
                    fl.swap(merged);
                }
            }
            // tiles of a row that a neighbouring slab reads go first: they are launched on their own, ahead of the halo push
            const bool hasLower = ctx->ring || ctx->g.rank > 0, hasUpper = ctx->ring || ctx->g.rank < ctx->g.nranks - 1;
            for(int k = 0; k < 3; ++k)
            {
                auto isB = [&](const TileRec& t) { return (hasLower && t.y == 1) || (hasUpper && t.y == ctx->ly - 2); };
                auto mid = std::stable_partition(lists[k].begin(), lists[k].end(), isB);
                ctx->nbound[fam][k] = (unsigned)(mid - lists[k].begin());
            }
            for(int k = 0; k < 3; ++k)
            {
                ctx->kstat[(fam == 0 ? K_E_FAST : K_H_FAST) + k].alg_bytes = listBytes[k];
                ctx->ntiles[fam][k] = (unsigned)lists[k].size();
                TileRec* d = nullptr;
                if((rc = dev_upload(ctx, &d, lists[k]))) return rc;
                ctx->d_tiles[fam][k] = d;
            }
        }
        cudaFree(d_sum);
    }
    // TFSF surfaces: pair lists and eps / mu lines to the device; no surface cell may lie inside the CPML
    if(!ctx->tfsf.empty())
    {
        int* d_terr = nullptr;
        if((rc = dev_alloc(ctx, &d_terr, 1))) return rc;
        for(TfsfDev& t : ctx->tfsf)
        {
            if((rc = dev_upload(ctx, &t.d_pairs_D, t.h_pairs_D))) return rc;
            if((rc = dev_upload(ctx, &t.d_pairs_U, t.h_pairs_U))) return rc;
            if(!t.h_ep_mu.empty() && (rc = dev_upload(ctx, &t.d_ep_mu, t.h_ep_mu))) return rc;
            for(int side = 0; side < 2; ++side)
            {
                const int np = side ? t.s.npairs_U : t.s.npairs_D;
                if(np == 0 || t.s.n == 0) continue;
                k_tfsf_check<<<(unsigned)std::min<long>(((long)np * t.s.n + 255) / 256, 1024), 256, 0, ctx->stream>>>(
                    side ? t.d_pairs_U : t.d_pairs_D, np, t.s.n, t.s.stride_main, ctx->d_info[t.s.comp], ctx->lx, ctx->px, (long)ctx->nlogical, d_terr);
            }
        }
        int h_terr = 0;
        CK(cudaMemcpyAsync(&h_terr, d_terr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_terr);
        if(h_terr & 4) return fail(ctx, CHIML_ERR_ARG, "TFSF surface: a main-grid index lies outside the grid");
        if(h_terr & 8) return fail(ctx, CHIML_ERR_UNSUPPORTED, "TFSF surface: a surface cell lies inside the CPML (the reference adds the incident field between the curl and the CPML terms)");
    }
    if((rc = build_tensor_maps(ctx))) return rc;
    // y-slab halo: flag words, push counters, the dense ghost row of node P_y (filled by the slab above)
    if(ctx->g.nranks > 1)
    {
        if((rc = dev_alloc_ipc(ctx, &ctx->d_flags, (size_t)HF_NFLAGS))) return rc;
        if((rc = dev_alloc(ctx, &ctx->d_push_counter, 16))) return rc;
        if(ctx->g.rank < ctx->g.nranks - 1)
            for(int p = 0; p < ctx->nordip; ++p)
                if((rc = dev_alloc_ipc(ctx, &ctx->d_oPy_ghost[p], (size_t)ctx->lx * ctx->lz))) return rc;
    }
    // emitters: constants, P boxes, SoA density state (rho_00 = weight of the level system, ML/density.hpp:57-60)
    for(EmitterDev& em : ctx->emitters)
    {
        em.group = emitter_group(em.d.nlevel);       // which density kernel, hence which state layout
        if((rc = dev_upload(ctx, &em.d_h0, em.h_h0))) return rc;
        if((rc = dev_upload(ctx, &em.d_mu, em.h_mu))) return rc;
        if((rc = dev_upload(ctx, &em.d_gam_ptr, em.h_gam_ptr))) return rc;
        if((rc = dev_upload(ctx, &em.d_gam_col, em.h_gam_col))) return rc;
        if((rc = dev_upload(ctx, &em.d_gam_val, em.h_gam_val))) return rc;
        if((rc = dev_upload(ctx, &em.d_loc, em.h_loc))) return rc;
        if((rc = dev_upload(ctx, &em.d_eps, em.h_eps))) return rc;
        for(int c = 0; c < 3; ++c) if((rc = dev_alloc_ipc(ctx, &em.d_P[c], em.pbox))) return rc;
        const size_t per = (size_t)em.d.nsys * em.n2 * 2 * (size_t)std::max(em.d.nemit, 1);
        std::vector<double> rho0(per, 0.0);
        for(int sy = 0; sy < em.d.nsys; ++sy)
            for(int e = 0; e < em.d.nemit; ++e)      // element (sys, e, k = 0), real part, in the layout of the kernel that will run
                rho0[em.group ? ((size_t)sy * em.d.nemit + e) * em.n2 * 2 : ((size_t)sy * em.n2 * 2) * em.d.nemit + e] = em.h_weight[sy];
        if((rc = dev_upload(ctx, &em.d_rho, rho0))) return rc;
        for(int k = 0; k < 4; ++k) if((rc = dev_alloc(ctx, &em.d_f[k], per))) return rc;
        const int epb = em.group ? 128 / em.group : 128;
        em.nblocks = (em.d.nemit + epb - 1) / epb;
        em.gam_maxrow = 0;
        for(int ii = 0; ii < em.n2; ++ii) em.gam_maxrow = std::max(em.gam_maxrow, em.h_gam_ptr[ii + 1] - em.h_gam_ptr[ii]);
        if((rc = dev_alloc(ctx, &em.d_pop_partial, (size_t)std::max(em.d.npop, 1) * std::max(em.nblocks, 1) * 2))) return rc;
        em.pop_cap = 4096;
        if((rc = dev_alloc(ctx, &em.d_pop, (size_t)std::max(em.d.npop, 1) * em.pop_cap * 2))) return rc;
        ctx->kstat[K_EMIT_DENSITY].alg_bytes += (double)em.d.nemit * (em.d.nsys * 96.0 * em.d.nlevel * em.d.nlevel + 96.0)   /* rho RW, 4 histories R, 1 W; 6 E reads, 3 P RMW */;
        ctx->kstat[K_EMIT_ADDP].alg_bytes += 0.0;   // its E read-modify-write is the field traffic already counted for the E half step
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for(DftDev& d : ctx->dfts)
    {
        if((rc = dev_upload(ctx, &d.d_lines, d.h_lines))) return rc;
        if((rc = dev_alloc(ctx, &d.d_re, d.acc_len))) return rc;
        if((rc = dev_alloc(ctx, &d.d_im, d.acc_len))) return rc;
    }
    // detectors: ring buffers sized on first use; sample at t = 0 (FDTD_MANAGER/parallelFDTDField.cpp:832-833)
    ctx->committed = true;
    for(size_t d = 0; d < ctx->detectors.size(); ++d)
    {
        DetectorDev& dt = ctx->detectors[d];
        dt.cap = std::min<size_t>(4096, std::max<size_t>(8, (64u << 20) / (dt.sample_len * sizeof(double))));   // grows on demand (reserve_rings)
        if((rc = dev_alloc(ctx, &dt.d_ring, dt.cap * dt.sample_len, false))) return rc;
        k_detector<<<(unsigned)std::min<size_t>((dt.sample_len + 255) / 256, 1024), 256, 0, ctx->stream>>>(
            ctx->d_field[dt.field], dt.loc[0], dt.loc[2], dt.loc[1], dt.sz[0], dt.sz[2], dt.sz[1], ctx->lz, ctx->px, dt.d_ring);
        ++ctx->launches;
        dt.count = 1;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    // host copies of the big lists are no longer needed
    for(auto& k : ctx->lists) for(auto& l : k) { l.runs.clear(); l.runs.shrink_to_fit(); }
    return CHIML_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------
// stepping
// ---------------------------------------------------------------------------------------------------
namespace {

void fill_step_args(ChimlCtx* ctx, bool isE, StepArgs& a)
{
    std::memset(&a, 0, sizeof(a));
    a.lx = ctx->lx; a.ly = ctx->ly; a.lz = ctx->lz; a.px = ctx->px;
    a.pml_on_D = isE ? ctx->g.pml_on_D : ctx->pml_on_B;       // (for the H family: the CPML acts on B)
    a.tmaps = reinterpret_cast<const unsigned char*>(ctx->d_tmaps);
    a.nsp_xmin = ctx->span_node.d_xmin; a.nsp_xmax = ctx->span_node.d_xmax; a.nsp_base = ctx->span_node.d_base;
    const int cur = ctx->pcur, prv = 1 - ctx->pcur;
    for(int i = 0; i < 3; ++i) a.fam[i] = ctx->d_field[(isE ? CHIML_HX : CHIML_EX) + i];
    for(int i = 0; i < 3; ++i)
    {
        const int comp = isE ? i : 3 + i;
        CompArgs& ca = a.c[i];
        if(!ctx->d_field[comp]) continue;
        ca.info = ctx->d_info[comp];
        ca.cls = ctx->d_cls[comp];
        ca.pf = ctx->d_pf[comp];
        ca.U = ctx->d_field[comp];
        ca.D = isE ? ctx->d_field[CHIML_DX + i] : (ctx->has_B ? ctx->d_field[CHIML_BX + i] : nullptr);
        const int base = isE ? CHIML_HX : CHIML_EX;
        ca.Vj = ctx->d_field[base + (i + 1) % 3];
        ca.Vk = ctx->d_field[base + (i + 2) % 3];
        int d[3];
        decode_offset(ctx, ctx->off_j[comp], d); ca.offJ = phys_offset(ctx, d);
        decode_offset(ctx, ctx->off_k[comp], d); ca.offK = phys_offset(ctx, d);
        for(int part = 0; part < 2; ++part)
        {
            const PmlPartDev& pp = ctx->pml[comp][part];
            PmlArgs& pa = ca.pml[part];
            pa.present = pp.present;
            if(!pp.present) continue;
            pa.F = pp.d_F; pa.b = pp.d_b; pa.c = pp.d_c; pa.cmap = pp.d_cmap; pa.psi = pp.d_psi;
            pa.Db = pp.Db; pa.psi_pitch = pp.psi_pitch; pa.axis = pp.axis; pa.nact = pp.nact; pa.has_psi = pp.has_psi;
        }
        if(ctx->has_chi)
        {
            const int cc = (isE ? 0 : 3) + i;
            for(int p = 0; p < MAX_CHI; ++p) { ca.chiCur[p] = ctx->d_chi[cc][p][cur]; ca.chiNew[p] = ctx->d_chi[cc][p][prv]; }
            ca.oppPrev = ctx->d_prev[(isE ? 3 : 0) + i];                 // E_c is driven by H_c and prevH_c, H_c by E_c and prevE_c
            int dd[3];
            decode_offset(ctx, ctx->chi_off[cc][0], dd); ca.chi_oi = phys_offset(ctx, dd);
            decode_offset(ctx, ctx->chi_off[cc][1], dd); ca.chi_oj = phys_offset(ctx, dd);
            decode_offset(ctx, ctx->chi_off[cc][2], dd); ca.chi_ok = phys_offset(ctx, dd);
        }
        if(!isE && ctx->has_B)
        {
            ca.sp_xmin = ctx->span[3 + i].d_xmin; ca.sp_base = ctx->span[3 + i].d_base;
            for(int p = 0; p < MAX_POLES; ++p) { ca.Pcur[p] = ctx->d_P[3 + i][p][cur]; ca.Pnew[p] = ctx->d_P[3 + i][p][prv]; }
        }
        if(isE)
        {
            ca.sp_xmin = ctx->span[i].d_xmin; ca.sp_base = ctx->span[i].d_base;
            for(int p = 0; p < MAX_POLES; ++p) { ca.Pcur[p] = ctx->d_P[i][p][cur]; ca.Pnew[p] = ctx->d_P[i][p][prv]; }
            ca.nordip = ctx->nordip;
            // after the node kernel of this step the new oriented-dipole P lives in buffer `prv`
            for(int p = 0; p < MAX_POLES; ++p) { ca.oP[p] = ctx->d_oP[i][p][prv]; ca.oPg[p] = i == 1 ? ctx->d_oPy_ghost[p] : nullptr; }
            decode_offset(ctx, ctx->ordip_off[i], d);
            ca.ord_dx = d[0]; ca.ord_dy = d[1]; ca.ord_dz = d[2];
            // orDipDtoUZ is bound for Ez when there is no Hz (FDTD_MANAGER/parallelFDTDField.cpp:293-296)
            ca.ord_zvariant = (i == 2 && !ctx->d_field[CHIML_HZ]) ? 1 : 0;
        }
    }
}

// Brackets one launch with CUDA events on the context's stream when kernel timing is on, and counts it.
struct LaunchScope
{
    ChimlCtx* ctx; int kind; cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t get(ChimlCtx* ctx)
    {
        cudaEvent_t e = nullptr;
        if(!ctx->ev_pool.empty()) { e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    LaunchScope(ChimlCtx* c, int k) : ctx(c), kind(k)
    {
        if(ctx->timing) { e0 = get(ctx); e1 = get(ctx); cudaEventRecord(e0, ctx->stream); }
    }
    ~LaunchScope()
    {
        ++ctx->launches;
        ++ctx->kstat[kind].launches;
        if(e0) { cudaEventRecord(e1, ctx->stream); ctx->ev_pending[kind].push_back({e0, e1}); }
    }
};

// part 0: every tile; part 1: the slab-boundary tiles (front of each list); part 2: the rest
template <bool IS_E, int MODE>
void launch_family_mode(ChimlCtx* ctx, const StepArgs& a, const dim3 block, int part)
{
    const int fam = IS_E ? 0 : 1;
    const int k0 = IS_E ? K_E_FAST : K_H_FAST;
    unsigned first[3], count[3];
    for(int k = 0; k < 3; ++k)
    {
        const unsigned n = ctx->ntiles[fam][k], nb = ctx->nbound[fam][k];
        first[k] = part == 2 ? nb : 0;
        count[k] = part == 0 ? n : (part == 1 ? nb : n - nb);
    }
    // 2-D grids: one-row tiles, ROWS_2D of them per block (chiml_update.cuh tile_of_thread)
    const bool twoD = MODE != CHIML_MODE_3D;
    auto nblk = [&](unsigned n) { return twoD ? (n + ROWS_2D - 1) / ROWS_2D : n; };
    const dim3 blk(block.x, twoD ? ROWS_2D : block.y, 1);
    if(count[0]) { LaunchScope ls(ctx, k0);     k_fast<IS_E, MODE><<<nblk(count[0]), blk, 0, ctx->stream>>>(a, (const TileRec*)ctx->d_tiles[fam][0] + first[0], count[0]); }
    if(count[1])
    {
        // 3-D: one block per (tile, z part, component) (chiml_update.cuh k_uniform); 2-D tiles are a single row
        LaunchScope ls(ctx, k0 + 1);
        constexpr unsigned zsplit = uniform_split<IS_E>();
        const TileRec* tl = (const TileRec*)ctx->d_tiles[fam][1] + first[1];
        if(twoD) k_uniform_rows<IS_E, MODE><<<nblk(count[1]), dim3(blk.x, blk.y, 3), 0, ctx->stream>>>(a, tl, count[1]);
        else k_uniform<IS_E, MODE><<<count[1] * zsplit * 3, dim3(block.x, block.y / zsplit, 1), 0, ctx->stream>>>(a, tl);
    }
    if(count[2])
    {
        LaunchScope ls(ctx, k0 + 2);
        if(!IS_E && ctx->has_B)       // magnetic-dispersive media: the H family with B targets, magnetic poles and B -> H
            k_general<IS_E, MODE, IS_E, true><<<nblk(count[2]), dim3(blk.x, blk.y, 1), 0, ctx->stream>>>(a, (const TileRec*)ctx->d_tiles[fam][2] + first[2], count[2]);
        else
            k_general<IS_E, MODE, IS_E><<<nblk(count[2]), dim3(blk.x, blk.y, IS_E ? 3 : 1), 0, ctx->stream>>>(a, (const TileRec*)ctx->d_tiles[fam][2] + first[2], count[2]);
    }
}
template <bool IS_E>
void launch_family(ChimlCtx* ctx, const StepArgs& a, const dim3 block, int part)
{
    if(ctx->g.mode == CHIML_MODE_3D)      launch_family_mode<IS_E, CHIML_MODE_3D>(ctx, a, block, part);
    else if(ctx->g.mode == CHIML_MODE_TE) launch_family_mode<IS_E, CHIML_MODE_TE>(ctx, a, block, part);
    else                                  launch_family_mode<IS_E, CHIML_MODE_TM>(ctx, a, block, part);
}

template <int N, int G>
void launch_density_g(ChimlCtx* ctx, const EmitArgs& ea)
{
    k_emit_density_g<N, G><<<(ea.nemit + 128 / G - 1) / (128 / G), 128, 0, ctx->stream>>>(ea);
}


// rows of [y0, y1) that are / are not slab-boundary rows (local row 1 with a slab below, row ly-2 with a slab above)
struct RowSeg { int y0, y1; };
int split_rows(const ChimlCtx* ctx, int y0, int y1, int part, RowSeg out[3])
{
    if(part == 0) { out[0] = {y0, y1}; return y1 > y0 ? 1 : 0; }
    const int bl = (ctx->ring || ctx->g.rank > 0) ? 1 : -1, bu = (ctx->ring || ctx->g.rank < ctx->g.nranks - 1) ? ctx->ly - 2 : -1;
    int n = 0;
    if(part == 1)
    {
        if(bl >= y0 && bl < y1) out[n++] = {bl, bl + 1};
        if(bu >= y0 && bu < y1 && bu != bl) out[n++] = {bu, bu + 1};
        return n;
    }
    int a = y0;
    for(int b : {bl, bu})
        if(b >= a && b < y1) { if(b > a) out[n++] = {a, b}; a = b + 1; }
    if(a < y1) out[n++] = {a, y1};
    return n;
}

void launch_addP(ChimlCtx* ctx, EmitterDev& em, int part)
{
    const bool threeD = ctx->lz > 1;
    AddPArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    for(int c = 0; c < 3; ++c) { pa.E[c] = ctx->d_field[c]; pa.P[c] = em.d_P[c]; }
    pa.eps = em.d_eps;
    for(int k = 0; k < 3; ++k) pa.box_lo[k] = em.d.box_lo[k];
    pa.nx = em.d.box_n[0] + 1; pa.ny = em.d.box_n[1] + 1; pa.nz = threeD ? em.d.box_n[2] + 1 : 1;
    pa.bx = em.d.box_n[0] + 2; pa.bz = em.pz; pa.zoff = threeD ? 1 : 0;
    pa.lz = ctx->lz; pa.px = ctx->px;
    RowSeg seg[3];
    // with several slabs the box may reach into a ghost row: that row belongs to the neighbour, which updates it and pushes it here
    // (adding P to it locally could land after the push and spoil it)
    const int r0 = ctx->g.nranks > 1 ? std::max(pa.box_lo[1], 1) : pa.box_lo[1];
    const int r1 = ctx->g.nranks > 1 ? std::min(pa.box_lo[1] + pa.ny, ctx->ly - 1) : pa.box_lo[1] + pa.ny;
    const int nseg = split_rows(ctx, r0, r1, part, seg);
    for(int i = 0; i < nseg; ++i)
    {
        pa.iy0 = seg[i].y0 - pa.box_lo[1]; pa.iy1 = seg[i].y1 - pa.box_lo[1];
        const long n = (long)pa.nx * (pa.iy1 - pa.iy0) * pa.nz;
        if(n <= 0) continue;
        LaunchScope ls(ctx, K_EMIT_ADDP);
        k_emit_addP<<<(unsigned)std::min<long>((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(pa);
    }
}

int launch_density_step(ChimlCtx* ctx, EmitterDev& em)
{
    const bool threeD = ctx->lz > 1;
    const int sample = (em.tstep % em.d.pop_every) == 0 && em.d.npop > 0;
    if(em.d.nemit > 0)
    {
        EmitArgs ea;
        std::memset(&ea, 0, sizeof(ea));
        for(int c = 0; c < 3; ++c) { ea.E[c] = ctx->d_field[c]; ea.P[c] = em.d_P[c]; ea.mu_present[c] = em.mu_present[c]; }
        ea.lz = ctx->lz; ea.px = ctx->px; ea.threeD = threeD ? 1 : 0; ea.tm = (!ctx->d_field[CHIML_EX]) ? 1 : 0;
        ea.nemit = em.d.nemit; ea.nsys = em.d.nsys; ea.nlevel = em.d.nlevel;
        for(int k = 0; k < 3; ++k) ea.box_lo[k] = em.d.box_lo[k];
        ea.bx = em.d.box_n[0] + 2; ea.bz = em.pz;
        ea.dt = em.d.dt; ea.inv_hbar = em.d.inv_hbar; ea.na = em.d.na;
        ea.h0 = em.d_h0; ea.mu = em.d_mu; ea.gam_ptr = em.d_gam_ptr; ea.gam_col = em.d_gam_col; ea.gam_val = em.d_gam_val;
        ea.loc = em.d_loc; ea.eps = em.d_eps;
        ea.rho = em.d_rho;
        for(int k = 0; k < 4; ++k) ea.f[k] = em.d_f[(em.fbase + k) % 4];
        ea.npop = em.d.npop; ea.sample = sample;
        for(int p = 0; p < em.d.npop; ++p) ea.pop_level[p] = em.h_pop_level[p];
        ea.pop_partial = em.d_pop_partial;
        {
            LaunchScope ls(ctx, K_EMIT_DENSITY);
            ea.gam_maxrow = em.gam_maxrow;
            switch(em.d.nlevel * (em.group ? 10 : 1))
            {
                case 2:  k_emit_density<2><<<em.nblocks, 128, 0, ctx->stream>>>(ea); break;
                case 3:  k_emit_density<3><<<em.nblocks, 128, 0, ctx->stream>>>(ea); break;
                case 4:  k_emit_density<4><<<em.nblocks, 128, 0, ctx->stream>>>(ea); break;
                case 5:  k_emit_density<5><<<em.nblocks, 128, 0, ctx->stream>>>(ea); break;
                case 6:  k_emit_density<6><<<em.nblocks, 128, 0, ctx->stream>>>(ea); break;
                case 20: launch_density_g<2, 4>(ctx, ea); break;
                case 30: launch_density_g<3, 16>(ctx, ea); break;
                case 40: launch_density_g<4, 16>(ctx, ea); break;
                default: launch_density_g<5, 32>(ctx, ea); break;
            }
        }
        em.fbase = (em.fbase + 3) % 4;   // the slot that held f_{n-3} now holds the new f_n
    }
    if(sample)
    {
        LaunchScope ls(ctx, K_EMIT_POP);
        k_emit_pop_reduce<<<1, 256, 0, ctx->stream>>>(em.d_pop_partial, em.d.nemit > 0 ? em.nblocks : 0, em.d.npop, 0.0, (double)em.d.npoints, em.d_pop, em.pop_cap, em.pop_n % em.pop_cap);
        ++em.pop_n;
    }
    ++em.tstep;
    return 0;
}

// part 0: all sources; part 1: H-field sources on slab-boundary rows (they must be in the row before it is pushed, and nothing of
// the H half step reads H); part 2: the rest -- E-field sources always come after the whole H half step, which reads E
void launch_sources(ChimlCtx* ctx, long long k, int nsrc, int part)
{
    for(int q = 0; q < nsrc; ++q)
    {
        const SourceDev& s = ctx->sources[q];
        const bool isH = s.field >= CHIML_HX && s.field <= CHIML_HZ;
        if(part == 1 && !isH) continue;
        RowSeg seg[3];
        const int nseg = split_rows(ctx, s.loc[1], s.loc[1] + s.sz[1], (part == 2 && !isH) ? 0 : part, seg);
        for(int i = 0; i < nseg; ++i)
        {
            const int sy = seg[i].y1 - seg[i].y0;
            const long n = (long)s.sz[0] * sy * s.sz[2];
            if(n <= 0) continue;
            LaunchScope ls(ctx, K_SOURCE);
            // (a source into H on a cell of upLorB_ is overwritten by B2H in the reference, which runs after the sources: skipped there)
            k_source<<<(unsigned)std::min<long>((n + 255) / 256, 2048), 256, 0, ctx->stream>>>(
                ctx->d_field[s.field], s.loc[0], s.loc[2], seg[i].y0, s.sz[0], s.sz[2], sy, ctx->lz, ctx->px, ctx->d_src_amp + k * nsrc + q,
                (isH && ctx->has_B) ? ctx->d_info[s.field] : nullptr);
        }
    }
}

// tfsf->updateFields() (step() item 5): the surface lists in the reference's order, grouped into waves of lists with distinct target arrays
// (two surfaces of one component share the cells of the edge where they meet: their additions keep the reference's order)
void launch_tfsf(ChimlCtx* ctx, long long k)
{
    if(ctx->tfsf.empty()) return;
    std::vector<TfsfEntry> order[9];                 // per target array (E, H, D components), in surface order: H surfaces first
    for(int pass = 0; pass < 2; ++pass)
        for(const TfsfDev& t : ctx->tfsf)
        {
            if((t.s.comp >= 3) != (pass == 0)) continue;
            for(int side = 0; side < 2; ++side)
            {
                const int np = side ? t.s.npairs_U : t.s.npairs_D;
                if(np == 0 || t.s.n == 0) continue;
                const int arr = side ? t.s.comp : CHIML_DX + t.s.comp;
                TfsfEntry e;
                e.target = ctx->d_field[arr]; e.pairs = side ? t.d_pairs_U : t.d_pairs_D; e.ep_mu = side ? t.d_ep_mu : nullptr;
                e.npairs = np; e.n = t.s.n; e.stride_incd = t.s.stride_incd; e.stride_main = t.s.stride_main; e.incd_offset = t.s.incd_offset;
                e.prefactor = t.s.prefactor;
                order[arr].push_back(e);
            }
        }
    size_t waves = 0;
    for(auto& o : order) waves = std::max(waves, o.size());
    for(size_t w = 0; w < waves; ++w)
    {
        TfsfArgs a;
        std::memset(&a, 0, sizeof(a));
        a.incd = ctx->d_tfsf_incd + (size_t)k * ctx->tfsf_per_step; a.lx = ctx->lx; a.px = ctx->px;
        long most = 1;
        for(auto& o : order) if(w < o.size()) { a.e[a.n++] = o[w]; most = std::max(most, (long)o[w].npairs * o[w].n); }
        LaunchScope ls(ctx, K_TFSF);
        k_tfsf<<<dim3((unsigned)std::min<long>((most + 255) / 256, 148 * 2), a.n, 1), 256, 0, ctx->stream>>>(a);
    }
}

// copy2PrevFields_ after updateChiH (E -> prevE, right after the H family, whose chiral magnetisation read them) and after updateChiE (H -> prevH)
void launch_prev_copy(ChimlCtx* ctx, bool copyE)
{
    if(!ctx->has_chi || ctx->n_prev_rows == 0) return;
    PrevCopyArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    for(int i = 0; i < 3; ++i) { pa.src[i] = ctx->d_field[(copyE ? CHIML_EX : CHIML_HX) + i]; pa.dst[i] = ctx->d_prev[(copyE ? 0 : 3) + i]; }
    pa.rows = ctx->d_prev_rows; pa.nrows = (unsigned)ctx->n_prev_rows; pa.lz = ctx->lz; pa.px = ctx->px;
    LaunchScope ls(ctx, K_PREV_COPY);
    k_prev_copy<<<pa.nrows, 64, 0, ctx->stream>>>(pa);
}

// applBCH_ / applBCE_ (step() items 9 and 17): periodic wrap copies of the three components of one family
void launch_wraps(ChimlCtx* ctx, bool isE)
{
    if(!ctx->periodic) return;
    WrapArgs wa;
    std::memset(&wa, 0, sizeof(wa));
    wa.lz = ctx->lz; wa.px = ctx->px; wa.ly = ctx->ly;
    long most = 0;
    for(int i = 0; i < 3; ++i)
    {
        const int comp = (isE ? 0 : 3) + i;
        if(!ctx->has_wrap[comp] || !ctx->d_field[comp]) continue;
        const ChimlWrap& w = ctx->wrap[comp];
        wa.f[wa.n] = ctx->d_field[comp]; wa.w[wa.n] = w; ++wa.n;
        // (a slab of several, ymax < 0: the owned rows 1 .. ly - 2 play the part of the rows 1 .. ymax - 1 of the shell, without the y faces)
        const long X = w.xmax + 1, Y = (w.ymax < 0 ? ctx->ly - 1 : w.ymax) + 1, Z = w.zmax - w.zmin + 2;
        most = std::max(most, w.zmin != 0 ? 2 * (X * Z + X * (Y - 2) + (Y - 2) * (Z - 2)) : 2L * (w.xmax - 1 + (w.ymax < 0 ? ctx->ly : w.ymax)));
    }
    if(wa.n == 0) return;
    LaunchScope ls(ctx, K_WRAP);
    k_wrap<<<dim3((unsigned)std::max<long>(1, std::min<long>((most + 255) / 256, 148 * 4)), wa.n, 1), 256, 0, ctx->stream>>>(wa);
}

// Bloch wrap copies of one family of a complex-field pair: phase table per component as the reference writes the arguments
// (UTIL/FDTD_up_eq.cpp:1248-1324: exp(cplx(0, +-k_x dx xmax +- k_y dy ymax +- k_z dz zmax)), summed left to right, x then y then z)
void launch_bloch(ChimlCtx* re, ChimlCtx* im, bool isE)
{
    BlochArgs ba;
    std::memset(&ba, 0, sizeof(ba));
    ba.lz = re->lz; ba.px = re->px;
    long most = 0;
    const double* k = re->k_point;
    const double dx = re->g.d[0], dy = re->g.d[1], dz = re->g.d[2];
    for(int i = 0; i < 3; ++i)
    {
        const int comp = (isE ? 0 : 3) + i;
        if(!re->has_wrap[comp] || !re->d_field[comp] || !im->d_field[comp]) continue;
        const ChimlWrap& w = re->wrap[comp];
        const int q = ba.n++;
        ba.fr[q] = re->d_field[comp]; ba.fi[q] = im->d_field[comp]; ba.w[q] = w;
        for(int cz = -1; cz <= 1; ++cz)
            for(int cy = -1; cy <= 1; ++cy)
                for(int cx = -1; cx <= 1; ++cx)
                {
                    double arg = 0.0;
                    bool first = true;
                    // the first wrapped axis enters as k * d * max or -1.0 * k * d * max, the following ones are added or subtracted
                    auto term = [&](int cdir, double kk, double dd, int mx) {
                        if(cdir == 0) return;
                        if(first) { arg = cdir > 0 ? kk * dd * mx : -1.0 * kk * dd * mx; first = false; }
                        else arg = cdir > 0 ? arg + kk * dd * mx : arg - kk * dd * mx;
                    };
                    term(cx, k[0], dx, w.xmax); term(cy, k[1], dy, w.ymax); term(cz, k[2], dz, w.zmax);
                    const std::complex<double> ph = std::exp(std::complex<double>(0.0, arg));
                    ba.ph[q][(cx + 1) + 3 * (cy + 1) + 9 * (cz + 1)] = make_double2(ph.real(), ph.imag());
                }
        const long X = w.xmax + 1, Y = w.ymax + 1, Z = w.zmax - w.zmin + 2;
        most = std::max(most, w.zmin != 0 ? 2 * (X * Z + X * (Y - 2) + (Y - 2) * (Z - 2)) : 2L * (w.xmax - 1 + w.ymax));
    }
    if(ba.n == 0) return;
    LaunchScope ls(re, K_WRAP_BLOCH);
    k_wrap_bloch<<<dim3((unsigned)std::max<long>(1, std::min<long>((most + 255) / 256, 148 * 4)), ba.n, 1), 256, 0, re->stream>>>(ba);
}

void launch_node_poles(ChimlCtx* ctx)
{
    if(!ctx->d_info_node) return;
    NodeArgs na;
    std::memset(&na, 0, sizeof(na));
    na.info = ctx->d_info_node; na.cls = ctx->d_cls_node;
    na.lx = ctx->lx; na.ly = ctx->ly; na.lz = ctx->lz; na.px = ctx->px;
    na.sp_xmin = ctx->span_node.d_xmin; na.sp_xmax = ctx->span_node.d_xmax; na.sp_base = ctx->span_node.d_base;
    na.rows = ctx->span_node.d_rows;
    const int cur = ctx->pcur, prv = 1 - ctx->pcur;
    for(int c = 0; c < 3; ++c)
    {
        na.E[c] = ctx->d_field[c];
        int d[3];
        decode_offset(ctx, ctx->node_off[c], d);
        na.eoff[c] = phys_offset(ctx, d);
        for(int p = 0; p < MAX_POLES; ++p) { na.Pcur[c][p] = ctx->d_oP[c][p][cur]; na.Pnew[c][p] = ctx->d_oP[c][p][prv]; na.dipg[c][p] = ctx->d_dipg[c][p]; }
    }
    LaunchScope ls(ctx, K_ORDIP_POLES);
    const dim3 ng((ctx->span_node.max_width + 255) / 256, ctx->span_node.nrows_used, 1);
    if(ng.y == 0) return;
    if(ctx->has_dipg) k_ordip_poles<true><<<ng, 256, 0, ctx->stream>>>(na);
    else              k_ordip_poles<false><<<ng, 256, 0, ctx->stream>>>(na);
}

// ---- halo helpers (chiml_halo.cuh) ------------------------------------------------------------------
// onHalo: the wait goes in front of a push on the halo stream (the compute stream keeps running)
void halo_wait(ChimlCtx* ctx, std::initializer_list<std::pair<int, long long>> flags, bool onHalo = false)
{
    HaloWaitArgs w;
    std::memset(&w, 0, sizeof(w));
    for(auto& f : flags)
        if(f.second > 0) { w.flag[w.n] = ctx->d_flags + f.first; w.value[w.n] = (int)f.second; ++w.n; }
    if(w.n == 0) return;
    w.error = ctx->d_flags + HF_ERROR;
    if(onHalo)
    {
        ++ctx->launches; ++ctx->kstat[K_HALO_WAIT].launches;
        k_halo_wait<<<1, 1, 0, ctx->hstream>>>(w);
        return;
    }
    LaunchScope ls(ctx, K_HALO_WAIT);
    k_halo_wait<<<1, 1, 0, ctx->stream>>>(w);
}

// the halo stream picks up after everything launched so far on the compute stream
void halo_fork(ChimlCtx* ctx)
{
    cudaEventRecord(ctx->ev_main, ctx->stream);
    cudaStreamWaitEvent(ctx->hstream, ctx->ev_main, 0);
}

void halo_push(ChimlCtx* ctx, const HaloPeer& peer, std::initializer_list<HaloSeg> segs, std::initializer_list<int> flags, long long value)
{
    HaloPushArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    long total = 0;
    for(auto& sg : segs) if(sg.src && sg.dst && sg.n > 0) { pa.seg[pa.nseg++] = sg; total += sg.n; }
    for(int f : flags) pa.peer_flag[pa.nflag++] = peer.flags + f;
    pa.value = (int)value;
    pa.counter = ctx->d_push_counter + (ctx->push_slot++ % 16);
    const unsigned blocks = (unsigned)std::max<long>(1, std::min<long>((total / 2 + 255) / 256, 64));
    ++ctx->launches; ++ctx->kstat[K_HALO_PUSH].launches;
    k_halo_push<<<blocks, 256, 0, ctx->hstream>>>(pa);
    ctx->push_pending = true;
}

int launch_step_slabs(ChimlCtx* ctx, long long k, int nsrc);

// section: 0 = the whole step; a complex-field pair interleaves its two parts: 1 = H half step without the wrap copies, 2 = E half step
// without them, 3 = the end of the step (buffer flip, detectors, running DFT)
int launch_step(ChimlCtx* ctx, long long k, int nsrc, int section = 0)
{
    // one block per tile of the compact lists built at commit
    const dim3 block(32, ctx->lz > 1 ? TILE_Z : 1, 1);
    StepArgs a;
    if(ctx->g.nranks > 1)
    {
        int rc = launch_step_slabs(ctx, k, nsrc);
        if(rc) return rc;
    }
    else
    {
        if(section == 0 || section == 1)
        {
            // H half step: updateH + updateHPML_ (step() items 4 and 6)
            fill_step_args(ctx, false, a);
            launch_family<false>(ctx, a, block, 0);
            launch_prev_copy(ctx, true);        // (the chiral magnetisation of this step has read E and prevE: E -> prevE)
            // TFSF surfaces (item 5): H after its curl, E / D before theirs and before the soft sources
            launch_tfsf(ctx, k);
            // sources (item 7): all sources, E and H alike, are injected here
            launch_sources(ctx, k, nsrc, 0);
        }
        // periodic boundaries of H (item 9)
        if(section == 0) launch_wraps(ctx, false);
        if(section == 0 || section == 2)
        {
            // oriented-dipole poles at the nodes (item 10, first loop)
            launch_node_poles(ctx);
            // E half step: isotropic poles, updateD/updateE, updateEPML_, D2E (items 10-15)
            fill_step_args(ctx, true, a);
            launch_family<true>(ctx, a, block, 0);
            launch_prev_copy(ctx, false);       // (the chiral polarisation of this step has read H and prevH: H -> prevH)
            // qe->addQE() for every emitter object (item 16)
            for(EmitterDev& em : ctx->emitters)
            {
                launch_addP(ctx, em, 0);
                int rc = launch_density_step(ctx, em);
                if(rc) return rc;
            }
        }
        // periodic boundaries of E (item 17)
        if(section == 0) launch_wraps(ctx, true);
        if(section == 1 || section == 2) return 0;
    }
    ctx->pcur = 1 - ctx->pcur;
    ++ctx->step_count;
    // detectors (item 18)
    for(auto& dt : ctx->detectors)
    {
        if(ctx->step_count % dt.every != 0) continue;
        {
            LaunchScope ls(ctx, K_DETECTOR);
            k_detector<<<(unsigned)std::min<size_t>((dt.sample_len + 255) / 256, 1024), 256, 0, ctx->stream>>>(
                ctx->d_field[dt.field], dt.loc[0], dt.loc[2], dt.loc[1], dt.sz[0], dt.sz[2], dt.sz[1], ctx->lz, ctx->px, dt.d_ring + (dt.count % dt.cap) * dt.sample_len);
        }
        ++dt.count;
    }
    // flux->fieldIn(tcur_) (item 18, :1300-1302)
    if(!ctx->dfts.empty())
    {
        size_t per_step = 0;
        std::vector<size_t> goff(ctx->dft_group_nfreq.size(), 0);
        for(size_t g = 0; g < ctx->dft_group_nfreq.size(); ++g) { goff[g] = per_step; per_step += 2 * (size_t)ctx->dft_group_nfreq[g]; }
        // every set that is due, DFT_BATCH of them per launch (the sets own their accumulators: no two touch the same entry)
        DftBatchArgs ba;
        ba.n = 0; ba.lx = ctx->lx; ba.px = ctx->px;
        size_t most = 0;
        auto flush = [&]() {
            if(ba.n == 0) return;
            LaunchScope ls(ctx, K_DFT);
            k_dft_batch<<<dim3((unsigned)std::max<size_t>(1, std::min<size_t>((most + 255) / 256, 148 * 2)), (unsigned)ba.n, 1), 256, 0, ctx->stream>>>(ba);
            ba.n = 0; most = 0;
        };
        for(DftDev& d : ctx->dfts)
        {
            if(ctx->step_count % d.every != 0 || d.nlines == 0) continue;
            DftBatchSet& o = ba.s[ba.n++];
            o.field = ctx->d_field[d.field]; o.lines = d.d_lines; o.nlines = d.nlines; o.npts = d.npts; o.stride = d.stride; o.nfreq = d.nfreq;
            o.tw = ctx->d_tw + (size_t)k * per_step + goff[d.group]; o.re = d.d_re; o.im = d.d_im;
            most = std::max(most, (size_t)d.nlines * (size_t)d.npts * d.nfreq);
            if(ba.n == DFT_BATCH) flush();
        }
        flush();
    }
    return 0;
}

// One time step of one y-slab of several (protocol and its proof of equivalence: chiml_b200/slab.py, tests/test_slab_gloo.py).
// kk = number of the step being taken, counted from 1: the value published in the neighbours' flags.
int launch_step_slabs(ChimlCtx* ctx, long long k, int nsrc)
{
    if(!ctx->halo_bound) return CHIML_ERR_STATE;
    const dim3 block(32, ctx->lz > 1 ? TILE_Z : 1, 1);
    const long long kk = ctx->step_count + 1;
    const bool lo = ctx->lower.present, up = ctx->upper.present;
    const long rowN = ctx->plane;                                    // one (x, z) plane of doubles, padded
    const long top = (long)(ctx->ly - 2) * ctx->plane, ghostTop = (long)(ctx->ly - 1) * ctx->plane;
    // a periodic run closes the slabs into a ring (chiml_b200/slab.py): every slab has both neighbours; the last slab's components that are one
    // row short in y (Hx, Hz, Ey) end at row ly - 3, and its row ly - 2 of Hx, Hz is the wrap image of slab 0's row 1, pushed across the seam.
    // Ey rows feed oriented-dipole nodes and emitters only, which a ring of slabs does not carry.
    const bool ring = ctx->ring, lastSlab = ctx->g.rank == ctx->g.nranks - 1, seamTop = ring && lastSlab, seamBottom = ring && ctx->g.rank == 0;
    // On a ring Ey rows are carried for emitters only.  The reference updates its emitters BEFORE it wraps E (step() items 16, 17): their
    // averages Ey[r], Ey[r - y] see the wrap rows of Ey as the step before left them, so the two seam rows of Ey (the last slab's row ly - 3 into
    // slab 0's row 0, slab 0's row 1 into the last slab's row ly - 2) travel at the END of the step; between the other slabs nothing changes.
    const bool haveEy = ctx->d_field[CHIML_EY] != nullptr && (!ring || !ctx->emitters.empty());
    const bool seamEy = ring && haveEy;
    const bool needEy = haveEy && !ring;
    StepArgs a;
    // pushes of the previous step read rows this step overwrites
    if(ctx->push_pending) { cudaStreamWaitEvent(ctx->stream, ctx->ev_push, 0); ctx->push_pending = false; }

    // ---- H half step: boundary rows first (they read the ghost E row the slab above pushed at the end of the last step)
    if(up) halo_wait(ctx, {{HF_E_FROM_UPPER, kk - 1}});
    fill_step_args(ctx, false, a);
    launch_family<false>(ctx, a, block, 1);
    launch_sources(ctx, k, nsrc, 1);
    auto pushHUp = [&](long srcRow) {
        halo_fork(ctx);
        const HaloPeer& p = ctx->upper;
        halo_push(ctx, p, {{ctx->d_field[CHIML_HX] ? ctx->d_field[CHIML_HX] + srcRow : nullptr, p.field[CHIML_HX], rowN},
                           {ctx->d_field[CHIML_HZ] ? ctx->d_field[CHIML_HZ] + srcRow : nullptr, p.field[CHIML_HZ], rowN}}, {HF_H_FROM_LOWER}, kk);
    };
    if(up && !seamTop) pushHUp(top);
    if(seamTop)
    {
        // the wrap row ly - 2 of Hx, Hz is free: the E half step of the step before has read it, and this step's own (discarded) update of its
        // cells -- the CPML lists of the reference cover that row, the wrap then overwrites it -- is through: slab 0 may push
        halo_fork(ctx);
        halo_push(ctx, ctx->upper, {}, {HF_SEAM_ACK}, kk);
    }
    if(seamBottom)
    {
        // slab 0's Hx, Hz row 1 into the wrap row ly - 2 of the last slab, once that slab has released the row for this step
        halo_fork(ctx);
        halo_wait(ctx, {{HF_SEAM_ACK, kk}}, true);
        const HaloPeer& p = ctx->lower;
        const long wrapRow = (long)(p.ly - 2) * ctx->plane;
        halo_push(ctx, p, {{ctx->d_field[CHIML_HX] ? ctx->d_field[CHIML_HX] + ctx->plane : nullptr, p.field[CHIML_HX] ? p.field[CHIML_HX] + wrapRow : nullptr, rowN},
                           {ctx->d_field[CHIML_HZ] ? ctx->d_field[CHIML_HZ] + ctx->plane : nullptr, p.field[CHIML_HZ] ? p.field[CHIML_HZ] + wrapRow : nullptr, rowN}},
                  {HF_H_FROM_UPPER}, kk);
    }
    launch_family<false>(ctx, a, block, 2);
    launch_sources(ctx, k, nsrc, 2);
    // (the last slab of a ring sends its top owned row of Hx, Hz, row ly - 3: an interior row, ready only now)
    if(seamTop) pushHUp(top - ctx->plane);
    // x / z wraps of the owned rows (slabs of a periodic run; nothing otherwise).  The last slab's wrap row arrives with the x / z ghost cells slab 0
    // had before ITS wrap: wait for the row, then the wrap below redoes them from the row's inner cells
    if(seamTop) halo_wait(ctx, {{HF_H_FROM_UPPER, kk}});
    launch_wraps(ctx, false);

    // ---- oriented-dipole poles at the nodes (read Ey of ghost row 0: pushed by the slab below after its last E half step)
    if(lo && needEy) halo_wait(ctx, {{HF_EY_FROM_LOWER, kk - 1}});
    launch_node_poles(ctx);
    if(lo && ctx->nordip > 0 && haveEy)
    {
        // every slab pushes, zeros where it holds no node cell: the count is the whole grid's (chiml_gpu_set_ordip_pole_count)
        halo_fork(ctx);
        NodePushArgs np;
        std::memset(&np, 0, sizeof(np));
        np.npoles = ctx->nordip;
        for(int p = 0; p < ctx->nordip; ++p) { np.pool[p] = ctx->d_oP[1][p][1 - ctx->pcur]; np.dst[p] = ctx->lower.oPy_ghost[p]; }
        np.sp_xmin = ctx->span_node.d_xmin; np.sp_xmax = ctx->span_node.d_xmax; np.sp_base = ctx->span_node.d_base;
        np.lx = ctx->lx; np.lz = ctx->lz;
        np.peer_flag = ctx->lower.flags + HF_OP_FROM_UPPER; np.value = (int)kk;
        np.counter = ctx->d_push_counter + (ctx->push_slot++ % 16);
        ++ctx->launches; ++ctx->kstat[K_HALO_PUSH].launches;
        k_halo_push_nodes<<<32, 256, 0, ctx->hstream>>>(np);
        ctx->push_pending = true;
    }

    // ---- E half step: boundary rows first
    bool recvQP = false;
    for(size_t q = 0; q < ctx->emitters.size(); ++q)
        if(up && haveEy && q < ctx->upper.emitPy.size() && ctx->upper.emitPy[q]) recvQP = true;
    halo_wait(ctx, {{HF_H_FROM_LOWER, lo ? kk : 0}, {HF_OP_FROM_UPPER, (up && ctx->nordip > 0 && haveEy) ? kk : 0},
                    {HF_QP_FROM_UPPER, recvQP ? kk - 1 : 0}});
    // (emitters on a ring: the last slab's wrap row of Ey must hold the row slab 0 sent at the end of the step before BEFORE this half step's own
    // writes to that row -- the reference's lists cover it -- so that the density update below reads what the reference reads)
    if(seamEy && seamTop) halo_wait(ctx, {{HF_EY_FROM_UPPER, kk - 1}});
    fill_step_args(ctx, true, a);
    launch_family<true>(ctx, a, block, 1);
    for(EmitterDev& em : ctx->emitters) launch_addP(ctx, em, 1);
    halo_fork(ctx);
    if(lo)
    {
        const HaloPeer& p = ctx->lower;
        const long pg = (long)(p.ly - 1) * ctx->plane;     // its upper ghost row
        halo_push(ctx, p, {{ctx->d_field[CHIML_EX] ? ctx->d_field[CHIML_EX] + ctx->plane : nullptr, p.field[CHIML_EX] ? p.field[CHIML_EX] + pg : nullptr, rowN},
                           {ctx->d_field[CHIML_EZ] ? ctx->d_field[CHIML_EZ] + ctx->plane : nullptr, p.field[CHIML_EZ] ? p.field[CHIML_EZ] + pg : nullptr, rowN}},
                  {HF_E_FROM_UPPER}, kk);
    }
    if(up && haveEy && !seamTop)
        halo_push(ctx, ctx->upper, {{ctx->d_field[CHIML_EY] + top, ctx->upper.field[CHIML_EY], rowN}}, {HF_EY_FROM_LOWER}, kk);
    launch_family<true>(ctx, a, block, 2);
    for(EmitterDev& em : ctx->emitters) launch_addP(ctx, em, 2);

    // ---- emitter density update (averages Ey[r], Ey[r - y]: needs this step's Ey in ghost row 0; slab 0 of a ring: the wrap row of the step before)
    if(!ctx->emitters.empty())
    {
        if(lo && haveEy) halo_wait(ctx, {{HF_EY_FROM_LOWER, (seamEy && seamBottom) ? kk - 1 : kk}});
        for(EmitterDev& em : ctx->emitters)
        {
            int rc = launch_density_step(ctx, em);
            if(rc) return rc;
        }
        if(lo && haveEy)
        {
            std::vector<size_t> qs;
            for(size_t q = 0; q < ctx->emitters.size(); ++q)
                if(ctx->emitters[q].d.box_lo[1] == 0 && q < ctx->lower.emitPy.size() && ctx->lower.emitPy[q]) qs.push_back(q);
            if(!qs.empty()) halo_fork(ctx);
            for(size_t i = 0; i < qs.size(); ++i)
            {
                EmitterDev& em = ctx->emitters[qs[i]];
                const long rowP = (long)(em.d.box_n[0] + 2) * em.pz;
                double* dst = ctx->lower.emitPy[qs[i]] + (long)(ctx->lower.emit_bn1[qs[i]] + 1) * rowP;
                // the flag is published with the last set only: pushes run in order on the halo stream
                if(i + 1 == qs.size()) halo_push(ctx, ctx->lower, {{em.d_P[1] + rowP, dst, rowP}}, {HF_QP_FROM_UPPER}, kk);
                else                   halo_push(ctx, ctx->lower, {{em.d_P[1] + rowP, dst, rowP}}, {}, kk);
            }
        }
    }
    // periodic boundaries of E (item 17): after the emitters, as in the reference
    launch_wraps(ctx, true);
    if(seamEy && (seamTop || seamBottom))
    {
        // the seam rows of Ey: each side first releases the row the other writes (its emitters have read it), then waits for the other's release
        halo_fork(ctx);
        if(seamTop) halo_push(ctx, ctx->upper, {}, {HF_EY_ACK}, kk);
        if(seamBottom) halo_push(ctx, ctx->lower, {}, {HF_EY_ACK}, kk);
        // (a ring of two: both roles on one slab pair; the release flag of either side lives in the other's memory, one word each)
        halo_wait(ctx, {{HF_EY_ACK, kk}}, true);
        if(seamTop)
            halo_push(ctx, ctx->upper, {{ctx->d_field[CHIML_EY] + top - ctx->plane, ctx->upper.field[CHIML_EY], rowN}}, {HF_EY_FROM_LOWER}, kk);
        if(seamBottom)
            halo_push(ctx, ctx->lower, {{ctx->d_field[CHIML_EY] + ctx->plane, ctx->lower.field[CHIML_EY] ? ctx->lower.field[CHIML_EY] + (long)(ctx->lower.ly - 2) * ctx->plane : nullptr, rowN}},
                      {HF_EY_FROM_UPPER}, kk);
    }
    (void)ghostTop;
    if(ctx->push_pending) cudaEventRecord(ctx->ev_push, ctx->hstream);
    return 0;
}

// samples a detector with interval `every` takes during steps (sc, sc + n]
size_t samples_in(long long sc, long long n, int every) { return (size_t)((sc + n) / every - sc / every); }

// ---- 2-D grids: all n steps in one cooperative launch (chiml_persist.cuh) ------------------------------------------------------
template <int MODE> int persist_occupancy(int* perSM)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(perSM, k_steps_2d<MODE>, 256, 0) == cudaSuccess ? 0 : 1;
}

bool persist_eligible(ChimlCtx* ctx)
{
    if(ctx->persist_mode == 0 || std::getenv("CHIML_B200_NO_PERSIST")) return false;
    if(ctx->imag || ctx->is_imag_part) return false;
    if(ctx->has_B) return false;              // B / M cells take the launch path
    if(!ctx->tfsf.empty()) return false;      // the surface waves are separate launches
    if(ctx->g.mode == CHIML_MODE_3D || ctx->g.nranks > 1 || !ctx->emitters.empty() || ctx->d_info_node) return false;
    // the resident grid holds 16-32 warps per SM, the launch-per-phase kernels 64: on a large 2-D grid the launches cost less than the lost
    // parallelism (C2, 2048^2: 0.142 ms per step by launches, 0.159 in one launch; C1, 512^2: 0.066 against 0.039); chiml_gpu_set_persistent(1) forces it
    if(ctx->persist_mode < 0 && (long)ctx->lx * ctx->ly > 1500000L) return false;
    if((int)ctx->detectors.size() > P2D_MAX_DET || (int)ctx->dfts.size() > P2D_MAX_DFT) return false;
    // sources are injected by one grid-wide pass: two boxes on the same field must not overlap (the launch path adds them one after the other)
    for(size_t i = 0; i < ctx->sources.size(); ++i)
        for(size_t j = i + 1; j < ctx->sources.size(); ++j)
        {
            const SourceDev &p = ctx->sources[i], &q = ctx->sources[j];
            bool apart = p.field != q.field;
            for(int k = 0; k < 3; ++k) if(p.loc[k] + p.sz[k] <= q.loc[k] || q.loc[k] + q.sz[k] <= p.loc[k]) apart = true;
            if(!apart) return false;
        }
    if(ctx->persist_blocks < 0)
    {
        ctx->persist_blocks = 0;
        int coop = 0, sms = 0, perSM = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        const int bad = ctx->g.mode == CHIML_MODE_TE ? persist_occupancy<CHIML_MODE_TE>(&perSM) : persist_occupancy<CHIML_MODE_TM>(&perSM);
        int cap = 4;                              // measured on C1 / C2 (profiles/README.md): more resident blocks only lengthen the grid barriers
        if(const char* ev = std::getenv("CHIML_B200_PERSIST_PER_SM")) cap = std::max(1, std::atoi(ev));
        if(coop && !bad && perSM > 0) ctx->persist_blocks = sms * std::min(perSM, cap);
        cudaGetLastError();
    }
    return ctx->persist_blocks > 0;
}

int launch_steps_2d(ChimlCtx* ctx, int n, int nsrc)
{
    Persist2DArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    if(!ctx->d_persist_sa)
    {
        // the argument blocks of the two half steps, the E one for either parity of the pole buffers: uploaded once
        StepArgs sa[3];
        const int keep = ctx->pcur;
        fill_step_args(ctx, false, sa[0]);
        ctx->pcur = 0; fill_step_args(ctx, true, sa[1]);
        ctx->pcur = 1; fill_step_args(ctx, true, sa[2]);
        ctx->pcur = keep;
        CK(cudaMalloc(&ctx->d_persist_sa, sizeof(sa))); ctx->dev_bytes += sizeof(sa);
        // ON THE CONTEXT'S STREAM: a synchronous cudaMemcpy from pageable memory returns once the data is staged, the DMA follows on the
        // legacy stream -- which a non-blocking stream does not wait for; with a busy copy engine the kernel below then read the argument
        // blocks of whichever simulation owned this allocation before
        CK(cudaMemcpyAsync(ctx->d_persist_sa, sa, sizeof(sa), cudaMemcpyHostToDevice, ctx->stream));
    }
    pa.sa = reinterpret_cast<const StepArgs*>(ctx->d_persist_sa);
    pa.pcur0 = ctx->pcur;
    for(int fam = 0; fam < 2; ++fam)
        for(int k = 0; k < 3; ++k) { pa.tiles[fam][k] = (const TileRec*)ctx->d_tiles[fam][k]; pa.ntiles[fam][k] = ctx->ntiles[fam][k]; }
    pa.nsteps = n; pa.nsrc = nsrc; pa.step0 = ctx->step_count;
    for(int q = 0; q < nsrc; ++q) pa.src[q] = ctx->sources[q];
    pa.src_amp = ctx->d_src_amp;
    for(int f = 0; f < CHIML_NFIELDS; ++f) pa.field[f] = ctx->d_field[f];
    pa.ndet = (int)ctx->detectors.size();
    for(int d = 0; d < pa.ndet; ++d)
    {
        const DetectorDev& dt = ctx->detectors[d];
        P2DDetector& o = pa.det[d];
        o.field = dt.field; o.every = dt.every; o.ring = dt.d_ring; o.cap = dt.cap; o.count0 = dt.count; o.sample_len = dt.sample_len;
        for(int k = 0; k < 3; ++k) { o.loc[k] = dt.loc[k]; o.sz[k] = dt.sz[k]; }
    }
    pa.ndft = (int)ctx->dfts.size();
    size_t per_step = 0;
    std::vector<size_t> goff(ctx->dft_group_nfreq.size(), 0);
    for(size_t g = 0; g < ctx->dft_group_nfreq.size(); ++g) { goff[g] = per_step; per_step += 2 * (size_t)ctx->dft_group_nfreq[g]; }
    for(int q = 0; q < pa.ndft; ++q)
    {
        const DftDev& d = ctx->dfts[q];
        P2DDft& o = pa.dft[q];
        o.field = d.field; o.group = d.group; o.every = d.every; o.nfreq = d.nfreq; o.npts = d.npts; o.stride = d.stride;
        o.lines = d.d_lines; o.nlines = d.nlines; o.re = d.d_re; o.im = d.d_im; o.tw_off = goff[d.group];
    }
    pa.tw = ctx->d_tw; pa.tw_per_step = per_step;
    pa.lx = ctx->lx; pa.lz = ctx->lz; pa.px = ctx->px;
    pa.periodic = ctx->periodic ? 1 : 0;
    for(int c = 0; c < 6; ++c) { pa.wrap[c] = ctx->wrap[c]; pa.has_wrap[c] = ctx->has_wrap[c] && ctx->d_field[c] ? 1 : 0; }
    // no more blocks than there are work items in the largest phase (8 warps per block, one item per warp at a time)
    unsigned items = 1;
    for(int fam = 0; fam < 2; ++fam) items = std::max(items, ctx->ntiles[fam][0] + 3 * ctx->ntiles[fam][1] + 3 * ctx->ntiles[fam][2]);
    const unsigned blocks = std::max(1u, std::min((unsigned)ctx->persist_blocks, (items + 7) / 8));
    void* args[] = {&pa};
    {
        LaunchScope ls(ctx, K_STEPS_2D);
        const void* fn = ctx->g.mode == CHIML_MODE_TE ? (const void*)k_steps_2d<CHIML_MODE_TE> : (const void*)k_steps_2d<CHIML_MODE_TM>;
        CK(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(32, 8, 1), args, 0, ctx->stream));
    }
    // host bookkeeping of what the launch does on the device
    for(DetectorDev& dt : ctx->detectors) dt.count += samples_in(ctx->step_count, n, dt.every);
    ctx->step_count += n;
    if(n & 1) ctx->pcur = 1 - ctx->pcur;
    return 0;
}


// Makes room in the detector and population rings for the samples of the next n steps, so that the step loop itself never allocates
// or synchronises.  A ring grows only when the host has not consumed what it holds (chiml_gpu_consume_detector / _population).
int reserve_rings(ChimlCtx* ctx, long long n)
{
    for(DetectorDev& dt : ctx->detectors)
    {
        const size_t need = dt.count - dt.base + samples_in(ctx->step_count, n, dt.every);
        if(need <= dt.cap) continue;
        const size_t ncap = std::max(need, 2 * dt.cap);
        double* bigger = nullptr;
        CK(cudaMalloc((void**)&bigger, ncap * dt.sample_len * sizeof(double)));
        for(size_t s = dt.base; s < dt.count; )      // unwrap: sample s moves from slot s % cap to slot s % ncap, in contiguous pieces
        {
            const size_t run = std::min({dt.count - s, dt.cap - s % dt.cap, ncap - s % ncap});
            CK(cudaMemcpyAsync(bigger + (s % ncap) * dt.sample_len, dt.d_ring + (s % dt.cap) * dt.sample_len, run * dt.sample_len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            s += run;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(dt.d_ring);
        ctx->dev_bytes += (ncap - dt.cap) * dt.sample_len * sizeof(double);
        dt.d_ring = bigger; dt.cap = ncap;
    }
    for(EmitterDev& em : ctx->emitters)
    {
        if(em.d.npop == 0) continue;
        // sampled on the steps whose number, counted from 0, is a multiple of the interval: multiples of e in [tstep, tstep + n)
        const long long e = em.d.pop_every;
        const size_t need = em.pop_n - em.pop_base + (size_t)((em.tstep + n + e - 1) / e - (em.tstep + e - 1) / e);
        if(need <= em.pop_cap) continue;
        const size_t ncap = std::max(need, 2 * em.pop_cap);
        double* bigger = nullptr;
        CK(cudaMalloc((void**)&bigger, (size_t)em.d.npop * ncap * 2 * sizeof(double)));
        for(int p = 0; p < em.d.npop; ++p)
            for(size_t s = em.pop_base; s < em.pop_n; )
            {
                const size_t run = std::min({em.pop_n - s, em.pop_cap - s % em.pop_cap, ncap - s % ncap});
                CK(cudaMemcpyAsync(bigger + ((size_t)p * ncap + s % ncap) * 2, em.d_pop + ((size_t)p * em.pop_cap + s % em.pop_cap) * 2, run * 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                s += run;
            }
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(em.d_pop);
        ctx->dev_bytes += (size_t)em.d.npop * (ncap - em.pop_cap) * 2 * sizeof(double);
        em.d_pop = bigger; em.pop_cap = ncap;
    }
    return 0;
}

int upload_src_amp(ChimlCtx* ctx, int n, const double* src_amp)
{
    const int nsrc = (int)ctx->sources.size();
    if(nsrc == 0) return 0;
    if(!src_amp) return fail(ctx, CHIML_ERR_ARG, "sources registered but no amplitudes given");
    const size_t need = (size_t)n * nsrc;
    if(need > ctx->src_amp_cap)
    {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_src_amp);
        CK(cudaMalloc((void**)&ctx->d_src_amp, need * sizeof(double)));
        ctx->src_amp_cap = need;
    }
    if(need) CK(cudaMemcpyAsync(ctx->d_src_amp, src_amp, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// n steps of a complex-field pair: the two parts take every half step one after the other on one stream, the Bloch wrap copies couple them
int step_n_pair(ChimlCtx* re, int n, const double* amp_re, const double* amp_im)
{
    if(!re) return CHIML_ERR_ARG;
    ChimlCtx* const ctx = re;           // (the CK macro reports into `ctx`)
    ChimlCtx* im = re->imag;
    if(!im) return fail(re, CHIML_ERR_STATE, "step_n_cplx: no imaginary part is bound (chiml_gpu_bind_imag)");
    if(n < 0) return fail(re, CHIML_ERR_ARG, "negative step count");
    CK(cudaSetDevice(re->device));
    int rc;
    if((rc = reserve_rings(re, n)) || (rc = reserve_rings(im, n))) return rc;
    if((rc = upload_src_amp(re, n, amp_re))) return rc;
    if((rc = upload_src_amp(im, n, amp_im))) return fail(re, rc, "step_n_cplx: amplitudes of the imaginary part: " + im->err);
    const int nsrc = (int)re->sources.size();
    for(int k = 0; k < n; ++k)
    {
        if((rc = launch_step(re, k, nsrc, 1)) || (rc = launch_step(im, k, nsrc, 1))) return fail(re, rc, "launch failed");
        launch_bloch(re, im, false);
        if((rc = launch_step(re, k, nsrc, 2)) || (rc = launch_step(im, k, nsrc, 2))) return fail(re, rc, "launch failed");
        launch_bloch(re, im, true);
        if((rc = launch_step(re, k, nsrc, 3)) || (rc = launch_step(im, k, nsrc, 3))) return fail(re, rc, "launch failed");
    }
    CK(cudaGetLastError());
    return CHIML_OK;
}

int step_n_impl(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles = nullptr, const double* incd = nullptr, size_t incd_per_step = 0)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "step before commit");
    if(n < 0) return fail(ctx, CHIML_ERR_ARG, "negative step count");
    if(ctx->imag || ctx->is_imag_part) return fail(ctx, CHIML_ERR_STATE, "this context is one part of a complex-field pair: step it with chiml_gpu_step_n_cplx on the real part");
    CK(cudaSetDevice(ctx->device));
    if(!ctx->tfsf.empty())
    {
        if(!incd) return fail(ctx, CHIML_ERR_ARG, "TFSF surfaces are registered: step with chiml_gpu_step_n_tfsf and the incident lines");
        for(const TfsfDev& t : ctx->tfsf)
            if((size_t)t.s.incd_offset + (size_t)t.s.incd_len > incd_per_step) return fail(ctx, CHIML_ERR_ARG, "step_n_tfsf: a surface's incident line does not fit the per-step table");
        const size_t need = incd_per_step * (size_t)n;
        if(need > ctx->tfsf_incd_cap)
        {
            CK(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_tfsf_incd);
            CK(cudaMalloc((void**)&ctx->d_tfsf_incd, std::max<size_t>(need, 1) * sizeof(double)));
            ctx->tfsf_incd_cap = need;
        }
        if(need) CK(cudaMemcpyAsync(ctx->d_tfsf_incd, incd, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->tfsf_per_step = incd_per_step;
    }
    if(!ctx->dfts.empty())
    {
        if(!twiddles) return fail(ctx, CHIML_ERR_ARG, "running-DFT sets are registered: step with chiml_gpu_step_n_dft and the twiddle factors");
        size_t per_step = 0;
        for(int nf : ctx->dft_group_nfreq) per_step += 2 * (size_t)nf;
        const size_t need = per_step * (size_t)n;
        if(need > ctx->tw_cap)
        {
            CK(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_tw);
            CK(cudaMalloc((void**)&ctx->d_tw, std::max<size_t>(need, 1) * sizeof(double)));
            ctx->tw_cap = need;
        }
        if(need) CK(cudaMemcpyAsync(ctx->d_tw, twiddles, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    { int rc = reserve_rings(ctx, n); if(rc) return rc; }
    const int nsrc = (int)ctx->sources.size();
    { int rc = upload_src_amp(ctx, n, src_amp); if(rc) return rc; }
    // (a call of one or two steps costs the same either way: seven short launches against one cooperative launch)
    if(n > 2 && persist_eligible(ctx))
    {
        int rc = launch_steps_2d(ctx, n, nsrc);
        if(rc) return rc;
    }
    else
        for(int k = 0; k < n; ++k)
        {
            int rc = launch_step(ctx, k, nsrc);
            if(rc) return fail(ctx, rc, "launch failed");
        }
    CK(cudaGetLastError());
    return CHIML_OK;
}

} // namespace

extern "C" {

int chiml_gpu_step_n(ChimlCtx* ctx, int n, const double* src_amp) { return step_n_impl(ctx, n, src_amp); }
int chiml_gpu_step_n_dft(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles) { return step_n_impl(ctx, n, src_amp, twiddles); }
int chiml_gpu_step_n_cplx(ChimlCtx* re, int n, const double* src_amp_re, const double* src_amp_im) { return step_n_pair(re, n, src_amp_re, src_amp_im); }

int chiml_gpu_bind_imag(ChimlCtx* re, ChimlCtx* im, const double* k_point)
{
    if(!re || !im || re == im) return CHIML_ERR_ARG;
    if(!re->committed || !im->committed) return fail(re, CHIML_ERR_STATE, "bind_imag: both parts must be committed");
    if(re->imag || re->is_imag_part || im->imag || im->is_imag_part) return fail(re, CHIML_ERR_STATE, "bind_imag: a context is already part of a pair");
    if(!k_point) return fail(re, CHIML_ERR_ARG, "bind_imag: k-point");
    if(re->device != im->device || re->lx != im->lx || re->ly != im->ly || re->lz != im->lz || re->g.mode != im->g.mode || re->g.has_D != im->g.has_D)
        return fail(re, CHIML_ERR_ARG, "bind_imag: the two parts must be set up from the same lists on the same device");
    if(re->g.nranks > 1 || im->g.nranks > 1) return fail(re, CHIML_ERR_UNSUPPORTED, "complex fields are covered for single-slab runs only");
    if(!re->periodic || !im->periodic) return fail(re, CHIML_ERR_ARG, "bind_imag: complex fields belong to periodic runs: call chiml_gpu_set_periodic on both parts");
    for(int c = 0; c < 6; ++c)
        if(re->has_wrap[c] != im->has_wrap[c] || (re->has_wrap[c] && std::memcmp(&re->wrap[c], &im->wrap[c], sizeof(ChimlWrap)) != 0))
            return fail(re, CHIML_ERR_ARG, "bind_imag: the two parts have different wrap descriptions");
    for(ChimlCtx* c : {re, im})
        if(!c->emitters.empty() || !c->dfts.empty() || !c->tfsf.empty())
            return fail(re, CHIML_ERR_UNSUPPORTED, "complex fields with emitters / running-DFT sets / TFSF surfaces are outside the covered hot path");
    if(re->sources.size() != im->sources.size()) return fail(re, CHIML_ERR_ARG, "bind_imag: the two parts have different sources");
    ChimlCtx* const ctx = re;           // (the CK macro reports into `ctx`)
    CK(cudaSetDevice(re->device));
    CK(cudaStreamSynchronize(im->stream));
    CK(cudaStreamSynchronize(re->stream));
    cudaStreamDestroy(im->stream);
    im->stream = re->stream;                       // one stream: the halves of the two parts and the wrap copies are ordered by it
    im->is_imag_part = true;
    im->real_part = re;
    re->imag = im;
    for(int k = 0; k < 3; ++k) re->k_point[k] = im->k_point[k] = k_point[k];
    return 0;
}

int chiml_gpu_step_n_tfsf(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles, const double* incd, size_t incd_per_step)
{ return step_n_impl(ctx, n, src_amp, twiddles, incd, incd_per_step); }

int chiml_gpu_download_dft(ChimlCtx* ctx, int slot, double* re, double* im)
{
    if(!ctx || !re || !im) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "download_dft before commit");
    if(slot < 0 || slot >= (int)ctx->dfts.size()) return fail(ctx, CHIML_ERR_ARG, "download_dft: bad slot");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const DftDev& d = ctx->dfts[slot];
    if(d.acc_len)
    {
        CK(cudaMemcpy(re, d.d_re, d.acc_len * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(im, d.d_im, d.acc_len * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return CHIML_OK;
}

int chiml_gpu_sync(ChimlCtx* ctx)
{
    if(!ctx) return CHIML_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->hstream));
    if(ctx->d_flags)
    {
        int err = 0;
        CK(cudaMemcpy(&err, ctx->d_flags + HF_ERROR, sizeof(int), cudaMemcpyDeviceToHost));
        if(err) return fail(ctx, CHIML_ERR_STATE, "a neighbouring slab did not deliver its ghost row within 30 s (halo wait timed out)");
    }
    return CHIML_OK;
}

// ---- y-slab binding ------------------------------------------------------------------------------
namespace {
struct HaloBlobHdr
{
    uint32_t magic; int32_t rank, nranks, lx, ly, lz; int64_t px; int64_t guard; int32_t nordip, nsets; int32_t has_field[6];
};
struct HaloBlobSet { cudaIpcMemHandle_t h; int32_t has, box_lo1, box_n1, rowlen, object, box_lo0, box_lo2, box_n0, box_n2; };
// the set of the neighbouring slab that belongs to the same emitter object: same object index, same x / z box
inline bool same_object(const HaloBlobSet& s, const chiml::EmitterDev& em)
{
    return s.object == em.d.object && s.box_lo0 == em.d.box_lo[0] && s.box_lo2 == em.d.box_lo[2] && s.box_n0 == em.d.box_n[0] && s.box_n2 == em.d.box_n[2];
}
constexpr uint32_t HALO_MAGIC = 0x4F4C4148u;   // "HALO"
}

int chiml_gpu_halo_export(ChimlCtx* ctx, void* blob, size_t cap, size_t* size)
{
    if(!ctx || !size) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "halo_export before commit");
    if(ctx->g.nranks <= 1) return fail(ctx, CHIML_ERR_ARG, "halo_export: the grid description names a single slab");
    CK(cudaSetDevice(ctx->device));
    const size_t need = sizeof(HaloBlobHdr) + (6 + 1 + MAX_POLES) * sizeof(cudaIpcMemHandle_t) + ctx->emitters.size() * sizeof(HaloBlobSet);
    *size = need;
    if(!blob) return CHIML_OK;
    if(cap < need) return fail(ctx, CHIML_ERR_ARG, "halo_export: buffer too small");
    std::vector<char> out(need, 0);
    HaloBlobHdr h{};
    h.magic = HALO_MAGIC; h.rank = ctx->g.rank; h.nranks = ctx->g.nranks; h.lx = ctx->lx; h.ly = ctx->ly; h.lz = ctx->lz; h.px = ctx->px;
    h.guard = (int64_t)ctx->guard; h.nordip = ctx->nordip; h.nsets = (int32_t)ctx->emitters.size();
    cudaIpcMemHandle_t* hs = reinterpret_cast<cudaIpcMemHandle_t*>(out.data() + sizeof(HaloBlobHdr));
    for(int f = 0; f < 6; ++f)
        if(ctx->d_field_base[f]) { h.has_field[f] = 1; CK(cudaIpcGetMemHandle(&hs[f], ctx->d_field_base[f])); }
    CK(cudaIpcGetMemHandle(&hs[6], ctx->d_flags));
    for(int p = 0; p < MAX_POLES; ++p)
        if(ctx->d_oPy_ghost[p]) CK(cudaIpcGetMemHandle(&hs[7 + p], ctx->d_oPy_ghost[p]));
    HaloBlobSet* sets = reinterpret_cast<HaloBlobSet*>(out.data() + sizeof(HaloBlobHdr) + (7 + MAX_POLES) * sizeof(cudaIpcMemHandle_t));
    for(size_t q = 0; q < ctx->emitters.size(); ++q)
    {
        const EmitterDev& em = ctx->emitters[q];
        sets[q].box_lo1 = em.d.box_lo[1]; sets[q].box_n1 = em.d.box_n[1]; sets[q].rowlen = (em.d.box_n[0] + 2) * em.pz;
        sets[q].object = em.d.object; sets[q].box_lo0 = em.d.box_lo[0]; sets[q].box_lo2 = em.d.box_lo[2]; sets[q].box_n0 = em.d.box_n[0]; sets[q].box_n2 = em.d.box_n[2];
        if(em.d_P[1] && ctx->d_field[CHIML_EY]) { sets[q].has = 1; CK(cudaIpcGetMemHandle(&sets[q].h, em.d_P[1])); }
    }
    std::memcpy(out.data(), &h, sizeof(h));
    std::memcpy(blob, out.data(), need);
    return CHIML_OK;
}

static int halo_open_peer(ChimlCtx* ctx, const void* blob, size_t size, bool isLower, HaloPeer& peer)
{
    if(size < sizeof(HaloBlobHdr)) return fail(ctx, CHIML_ERR_ARG, "halo_bind: truncated blob");
    HaloBlobHdr h;
    std::memcpy(&h, blob, sizeof(h));
    if(h.magic != HALO_MAGIC) return fail(ctx, CHIML_ERR_ARG, "halo_bind: not a halo blob");
    const int nb = ctx->g.rank + (isLower ? -1 : 1);
    if(h.nranks != ctx->g.nranks || h.rank != (ctx->ring ? (nb + ctx->g.nranks) % ctx->g.nranks : nb)) return fail(ctx, CHIML_ERR_ARG, "halo_bind: blob is not from the neighbouring slab");
    if(h.lx != ctx->lx || h.lz != ctx->lz || h.px != ctx->px) return fail(ctx, CHIML_ERR_ARG, "halo_bind: neighbour has different x / z extents");
    const size_t need = sizeof(HaloBlobHdr) + (7 + MAX_POLES) * sizeof(cudaIpcMemHandle_t) + (size_t)h.nsets * sizeof(HaloBlobSet);
    if(size < need) return fail(ctx, CHIML_ERR_ARG, "halo_bind: truncated blob");
    const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>((const char*)blob + sizeof(HaloBlobHdr));
    auto open = [&](const cudaIpcMemHandle_t& hh, void** out) -> int {
        const std::string key(reinterpret_cast<const char*>(&hh), sizeof(hh));
        auto it = ctx->ipc_cache.find(key);
        if(it != ctx->ipc_cache.end()) { *out = it->second; return 0; }
        CK(cudaIpcOpenMemHandle(out, hh, cudaIpcMemLazyEnablePeerAccess));
        peer.opened.push_back(*out);
        ctx->ipc_cache[key] = *out;
        return 0;
    };
    int rc;
    peer.ly = h.ly;
    for(int f = 0; f < 6; ++f)
    {
        if(!h.has_field[f]) continue;
        // only the arrays this slab writes into: E_x, E_z of the slab below; H_x, H_z, E_y of the slab above
        // (periodic ring: slab 0 also writes H_x, H_z row 1 into the wrap row of the last slab, which is the slab below it)
        const bool seam = ctx->ring && ctx->g.rank == 0 && isLower && (f == CHIML_HX || f == CHIML_HZ || f == CHIML_EY);
        const bool needIt = seam || (isLower ? (f == CHIML_EX || f == CHIML_EZ) : (f == CHIML_HX || f == CHIML_HZ || f == CHIML_EY));
        if(!needIt) continue;
        void* base = nullptr;
        if((rc = open(hs[f], &base))) return rc;
        peer.field[f] = reinterpret_cast<double*>(base) + h.guard;
    }
    { void* base = nullptr; if((rc = open(hs[6], &base))) return rc; peer.flags = reinterpret_cast<int*>(base); }
    // the reference does not wrap the emitters' polarisation boxes: across the seam of a periodic ring no emitter set is paired
    const bool acrossSeam = ctx->ring && ((isLower && ctx->g.rank == 0) || (!isLower && ctx->g.rank == ctx->g.nranks - 1));
    if(acrossSeam)
    {
        peer.emitPy.assign(ctx->emitters.size(), nullptr);
        peer.emit_bn1.assign(ctx->emitters.size(), 0);
    }
    else if(isLower)
    {
        if(h.nordip != ctx->nordip)
            return fail(ctx, CHIML_ERR_ARG, "halo_bind: neighbour counts a different number of oriented-dipole pole grids: give every slab the count of "
                                            "the whole grid with chiml_gpu_set_ordip_pole_count");
        for(int p = 0; p < h.nordip; ++p) { void* base = nullptr; if((rc = open(hs[7 + p], &base))) return rc; peer.oPy_ghost[p] = reinterpret_cast<double*>(base); }
        const HaloBlobSet* sets = reinterpret_cast<const HaloBlobSet*>((const char*)blob + sizeof(HaloBlobHdr) + (7 + MAX_POLES) * sizeof(cudaIpcMemHandle_t));
        peer.emitPy.assign(ctx->emitters.size(), nullptr);
        peer.emit_bn1.assign(ctx->emitters.size(), 0);
        // a set of ours whose first row is inside the object pairs with THE set of the same object below whose top rim is its ghost row
        // (same object index and x / z box; every set of the neighbour is paired at most once)
        std::vector<char> used((size_t)std::max(h.nsets, 0), 0);
        for(size_t q = 0; q < ctx->emitters.size(); ++q)
        {
            const EmitterDev& em = ctx->emitters[q];
            if(em.d.box_lo[1] != 0 || !ctx->d_field[CHIML_EY]) continue;   // only P_y couples across a slab boundary
            int match = -1, nmatch = 0;
            for(int j = 0; j < h.nsets; ++j)
                if(sets[j].has && !used[j] && same_object(sets[j], em) && sets[j].rowlen == (em.d.box_n[0] + 2) * em.pz && sets[j].box_lo1 + sets[j].box_n1 + 1 == h.ly - 1)
                { if(match < 0) match = j; ++nmatch; }
            if(match < 0) return fail(ctx, CHIML_ERR_ARG, "halo_bind: an emitter object reaches this slab's first row but the slab below has no matching set");
            if(nmatch > 1) return fail(ctx, CHIML_ERR_ARG, "halo_bind: several emitter sets of the slab below match one set of this slab: give every object its own ChimlEmitterDesc::object");
            used[match] = 1;
            void* base = nullptr;
            if((rc = open(sets[match].h, &base))) return rc;
            peer.emitPy[q] = reinterpret_cast<double*>(base);
            peer.emit_bn1[q] = sets[match].box_n1;
        }
    }
    else
    {
        const HaloBlobSet* sets = reinterpret_cast<const HaloBlobSet*>((const char*)blob + sizeof(HaloBlobHdr) + (7 + MAX_POLES) * sizeof(cudaIpcMemHandle_t));
        peer.emitPy.assign(ctx->emitters.size(), nullptr);   // non-null marks "the slab above pushes into this set's rim"
        for(size_t q = 0; q < ctx->emitters.size(); ++q)
        {
            const EmitterDev& em = ctx->emitters[q];
            if(em.d.box_lo[1] + em.d.box_n[1] + 1 != ctx->ly - 1 || !ctx->d_field[CHIML_EY]) continue;
            int nmatch = 0;
            for(int j = 0; j < h.nsets; ++j) if(sets[j].has && sets[j].box_lo1 == 0 && same_object(sets[j], em) && sets[j].rowlen == (em.d.box_n[0] + 2) * em.pz) ++nmatch;
            if(nmatch == 0) return fail(ctx, CHIML_ERR_ARG, "halo_bind: an emitter box ends in this slab's ghost row but the slab above has no matching set");
            if(nmatch > 1) return fail(ctx, CHIML_ERR_ARG, "halo_bind: several emitter sets of the slab above match one set of this slab: give every object its own ChimlEmitterDesc::object");
            peer.emitPy[q] = em.d_P[1];
        }
    }
    peer.present = true;
    return CHIML_OK;
}

int chiml_gpu_halo_bind(ChimlCtx* ctx, const void* lower_blob, size_t lower_size, const void* upper_blob, size_t upper_size)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "halo_bind before commit");
    if(ctx->halo_bound) return fail(ctx, CHIML_ERR_STATE, "halo_bind called twice");
    CK(cudaSetDevice(ctx->device));
    const bool needLower = ctx->ring || ctx->g.rank > 0, needUpper = ctx->ring || ctx->g.rank < ctx->g.nranks - 1;
    if(needLower != (lower_blob != nullptr) || needUpper != (upper_blob != nullptr))
        return fail(ctx, CHIML_ERR_ARG, "halo_bind: exactly the existing neighbours must be given (none below slab 0, none above the last slab; a periodic "
                                        "run closes the ring: slab nranks - 1 below slab 0, slab 0 above slab nranks - 1)");
    int rc;
    if(needLower && (rc = halo_open_peer(ctx, lower_blob, lower_size, true, ctx->lower))) return rc;
    if(needUpper && (rc = halo_open_peer(ctx, upper_blob, upper_size, false, ctx->upper))) return rc;
    ctx->halo_bound = true;
    return CHIML_OK;
}

int chiml_gpu_step_n_timed(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles, float* ms)
{
    if(!ctx || !ms) return CHIML_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = step_n_impl(ctx, n, src_amp, twiddles);
    if(rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return CHIML_OK;
}

int64_t chiml_gpu_launch_count(const ChimlCtx* ctx) { return ctx ? ctx->launches : 0; }

int chiml_gpu_set_kernel_timing(ChimlCtx* ctx, int on)
{
    if(!ctx) return CHIML_ERR_ARG;
    ctx->timing = on != 0;
    return CHIML_OK;
}

int chiml_gpu_n_kernel_kinds(void) { return K_NKINDS; }

int chiml_gpu_kernel_stat(ChimlCtx* ctx, int kind, ChimlKernelStat* out)
{
    static const char* names[K_NKINDS] = {"k_fast<E>", "k_uniform<E>", "k_general<E>", "k_fast<H>", "k_uniform<H>", "k_general<H>",
                                          "k_ordip_poles", "k_source", "k_detector", "k_emit_addP", "k_emit_density", "k_emit_pop_reduce", "k_halo_push", "k_halo_wait", "k_dft", "k_steps_2d", "k_wrap", "k_tfsf", "k_wrap_bloch", "k_prev_copy"};
    if(!ctx || !out || kind < 0 || kind >= K_NKINDS) return CHIML_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    KernelStat& ks = ctx->kstat[kind];
    for(auto& pr : ctx->ev_pending[kind])
    {
        float ms = 0.f;
        if(cudaEventElapsedTime(&ms, pr[0], pr[1]) == cudaSuccess) { ks.ms_total += ms; ++ks.timed; }
        ctx->ev_pool.push_back(pr[0]); ctx->ev_pool.push_back(pr[1]);
    }
    ctx->ev_pending[kind].clear();
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->name, sizeof(out->name), "%s", names[kind]);
    out->launches = ks.launches; out->timed_launches = ks.timed; out->ms_total = ks.ms_total;
    out->alg_bytes_per_step = ks.alg_bytes;
    const long long steps = ctx->step_count - ctx->stat_step0;
    out->alg_bytes_per_launch = (ks.launches > 0 && steps > 0) ? ks.alg_bytes * (double)steps / (double)ks.launches : ks.alg_bytes;
    return CHIML_OK;
}

int chiml_gpu_reset_kernel_stats(ChimlCtx* ctx)
{
    if(!ctx) return CHIML_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for(int k = 0; k < K_NKINDS; ++k)
    {
        for(auto& pr : ctx->ev_pending[k]) { ctx->ev_pool.push_back(pr[0]); ctx->ev_pool.push_back(pr[1]); }
        ctx->ev_pending[k].clear();
        ctx->kstat[k].launches = 0; ctx->kstat[k].ms_total = 0.0; ctx->kstat[k].timed = 0;
    }
    ctx->stat_step0 = ctx->step_count;
    return CHIML_OK;
}
size_t chiml_gpu_device_bytes(const ChimlCtx* ctx) { return ctx ? ctx->dev_bytes : 0; }

int chiml_gpu_upload_field(ChimlCtx* ctx, int field, const double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "upload before commit");
    if(field < 0 || field >= CHIML_NFIELDS || !ctx->d_field[field]) return fail(ctx, CHIML_ERR_ARG, "upload_field: field absent");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(ctx->d_field[field], ctx->px * sizeof(double), host, ctx->lx * sizeof(double), ctx->lx * sizeof(double),
                         (size_t)ctx->ly * ctx->lz, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CHIML_OK;
}

int chiml_gpu_download_field(ChimlCtx* ctx, int field, double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "download before commit");
    if(field < 0 || field >= CHIML_NFIELDS || !ctx->d_field[field]) return fail(ctx, CHIML_ERR_ARG, "download_field: field absent");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(host, ctx->lx * sizeof(double), ctx->d_field[field], ctx->px * sizeof(double), ctx->lx * sizeof(double),
                         (size_t)ctx->ly * ctx->lz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CHIML_OK;
}

static int pole_xfer(ChimlCtx* ctx, int comp, int pole, int prev, double* host, const double* hostIn)
{
    if(!ctx || (!host && !hostIn)) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "pole access before commit");
    if(comp < 0 || comp > 5 || pole < 0 || pole >= MAX_POLES) return fail(ctx, CHIML_ERR_ARG, "pole access: bad comp/pole");
    CK(cudaSetDevice(ctx->device));
    const SpanTable& sp = ctx->span[comp];
    double* pool = ctx->d_P[comp][pole][prev ? 1 - ctx->pcur : ctx->pcur];
    if(host) std::fill(host, host + ctx->nlogical, 0.0);
    if(!pool || sp.total == 0) return CHIML_OK;
    std::vector<double> tmp((size_t)sp.total);
    if(host) { CK(cudaMemcpyAsync(tmp.data(), pool, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    const size_t nrows = (size_t)ctx->ly * ctx->lz;
    for(size_t row = 0; row < nrows; ++row)
    {
        if(sp.h_xmin[row] < 0) continue;
        const int w = std::min(sp.h_xmax[row], ctx->lx - 1) - sp.h_xmin[row] + 1;     // the span's padding cell may lie beyond the logical row
        if(host) std::copy_n(&tmp[(size_t)sp.h_base[row]], w, host + row * ctx->lx + sp.h_xmin[row]);
        else     std::copy_n(hostIn + row * ctx->lx + sp.h_xmin[row], w, &tmp[(size_t)sp.h_base[row]]);
    }
    if(hostIn) { CK(cudaMemcpyAsync(pool, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    return CHIML_OK;
}

int chiml_gpu_download_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host) { return (comp < 0 || comp > 2) ? CHIML_ERR_ARG : pole_xfer(ctx, comp, pole, prev, host, nullptr); }
int chiml_gpu_upload_pole(ChimlCtx* ctx, int comp, int pole, int prev, const double* host) { return (comp < 0 || comp > 2) ? CHIML_ERR_ARG : pole_xfer(ctx, comp, pole, prev, nullptr, host); }
int chiml_gpu_download_chi_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "pole access before commit");
    if(comp < 0 || comp > 5 || pole < 0 || pole >= MAX_CHI) return fail(ctx, CHIML_ERR_ARG, "chiral pole access: bad comp/pole");
    CK(cudaSetDevice(ctx->device));
    const SpanTable& sp = ctx->span[comp];
    double* pool = ctx->d_chi[comp][pole][prev ? 1 - ctx->pcur : ctx->pcur];
    std::fill(host, host + ctx->nlogical, 0.0);
    if(!pool || sp.total == 0) return CHIML_OK;
    std::vector<double> tmp((size_t)sp.total);
    CK(cudaMemcpyAsync(tmp.data(), pool, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t nrows = (size_t)ctx->ly * ctx->lz;
    for(size_t row = 0; row < nrows; ++row)
    {
        if(sp.h_xmin[row] < 0) continue;
        const int w = std::min(sp.h_xmax[row], ctx->lx - 1) - sp.h_xmin[row] + 1;
        std::copy_n(&tmp[(size_t)sp.h_base[row]], w, host + row * ctx->lx + sp.h_xmin[row]);
    }
    return CHIML_OK;
}

int chiml_gpu_download_prev_field(ChimlCtx* ctx, int comp, double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "download_prev_field before commit");
    if(comp < 0 || comp > 5 || !ctx->d_prev[comp]) return fail(ctx, CHIML_ERR_ARG, "download_prev_field: no previous-field copy of this component (no chiral list)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(host, (size_t)ctx->lx * sizeof(double), ctx->d_prev[comp], (size_t)ctx->px * sizeof(double), (size_t)ctx->lx * sizeof(double),
                         (size_t)ctx->ly * ctx->lz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CHIML_OK;
}

int chiml_gpu_download_mag_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host) { return (comp < 0 || comp > 2) ? CHIML_ERR_ARG : pole_xfer(ctx, 3 + comp, pole, prev, host, nullptr); }

int chiml_gpu_download_ordip_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "pole access before commit");
    if(comp < 0 || comp > 2 || pole < 0 || pole >= MAX_POLES) return fail(ctx, CHIML_ERR_ARG, "pole access: bad comp/pole");
    CK(cudaSetDevice(ctx->device));
    const SpanTable& sp = ctx->span_node;
    double* pool = ctx->d_oP[comp][pole][prev ? 1 - ctx->pcur : ctx->pcur];
    std::fill(host, host + ctx->nlogical, 0.0);
    if(!pool || sp.total == 0) return CHIML_OK;
    std::vector<double> tmp((size_t)sp.total);
    CK(cudaMemcpyAsync(tmp.data(), pool, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t nrows = (size_t)ctx->ly * ctx->lz;
    for(size_t row = 0; row < nrows; ++row)
        if(sp.h_xmin[row] >= 0)
            std::copy_n(&tmp[(size_t)sp.h_base[row]], std::min(sp.h_xmax[row], ctx->lx - 1) - sp.h_xmin[row] + 1, host + row * ctx->lx + sp.h_xmin[row]);
    return CHIML_OK;
}

int chiml_gpu_download_psi(ChimlCtx* ctx, int comp, int part, double* host)
{
    if(!ctx || !host) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "psi access before commit");
    if(comp < 0 || comp > 5 || part < 0 || part > 1) return fail(ctx, CHIML_ERR_ARG, "download_psi: bad comp/part");
    CK(cudaSetDevice(ctx->device));
    const PmlPartDev& pp = ctx->pml[comp][part];
    std::fill(host, host + ctx->nlogical, 0.0);
    if(!pp.d_psi) return CHIML_OK;
    std::vector<double> tmp((size_t)pp.psi_count);
    CK(cudaMemcpyAsync(tmp.data(), pp.d_psi, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for(int y = 0; y < ctx->ly; ++y)
        for(int z = 0; z < ctx->lz; ++z)
            for(int x = 0; x < ctx->lx; ++x)
            {
                const int co = pp.axis == 0 ? x : (pp.axis == 1 ? y : z);
                const int cc = pp.h_cmap[co];
                if(cc < 0) continue;
                long ip;
                if(pp.axis == 0)      ip = cc + pp.psi_pitch * (z + (long)ctx->lz * y);
                else if(pp.axis == 1) ip = x + ctx->px * (z + (long)ctx->lz * cc);
                else                  ip = x + ctx->px * (cc + (long)pp.nact * y);
                host[x + (size_t)ctx->lx * (z + (size_t)ctx->lz * y)] = tmp[(size_t)ip];
            }
    return CHIML_OK;
}

} // extern "C"
// copies samples [first, first + n) of a detector ring to the host (the stream is idle: the callers synchronise first)
static int ring_read(ChimlCtx* ctx, const DetectorDev& dt, size_t first, size_t n, double* out)
{
    for(size_t s = first, done = 0; done < n; )
    {
        const size_t run = std::min(n - done, dt.cap - s % dt.cap);
        CK(cudaMemcpy(out + done * dt.sample_len, dt.d_ring + (s % dt.cap) * dt.sample_len, run * dt.sample_len * sizeof(double), cudaMemcpyDeviceToHost));
        s += run; done += run;
    }
    return CHIML_OK;
}
extern "C" {

int chiml_gpu_read_detector(ChimlCtx* ctx, int slot, double* out, size_t cap_samples, size_t* n_samples)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "read_detector before commit");
    if(slot < 0 || slot >= (int)ctx->detectors.size()) return fail(ctx, CHIML_ERR_ARG, "read_detector: bad slot");
    CK(cudaSetDevice(ctx->device));
    const DetectorDev& dt = ctx->detectors[slot];
    CK(cudaStreamSynchronize(ctx->stream));
    if(n_samples) *n_samples = dt.count - dt.base;
    const size_t n = std::min(cap_samples, dt.count - dt.base);
    return out && n ? ring_read(ctx, dt, dt.base, n, out) : CHIML_OK;
}

int chiml_gpu_consume_detector(ChimlCtx* ctx, int slot, size_t upto)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "consume_detector before commit");
    if(slot < 0 || slot >= (int)ctx->detectors.size()) return fail(ctx, CHIML_ERR_ARG, "consume_detector: bad slot");
    DetectorDev& dt = ctx->detectors[slot];
    dt.base = std::min(std::max(dt.base, upto), dt.count);
    return CHIML_OK;
}

int chiml_gpu_consume_population(ChimlCtx* ctx, int slot, size_t upto)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "consume_population before commit");
    if(slot < 0 || slot >= (int)ctx->emitters.size()) return fail(ctx, CHIML_ERR_ARG, "consume_population: bad slot");
    EmitterDev& em = ctx->emitters[slot];
    em.pop_base = std::min(std::max(em.pop_base, upto), em.pop_n);
    return CHIML_OK;
}

int chiml_gpu_reserve_steps(ChimlCtx* ctx, long long n_steps)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "reserve_steps before commit");
    if(n_steps < 0) return fail(ctx, CHIML_ERR_ARG, "reserve_steps: negative step count");
    CK(cudaSetDevice(ctx->device));
    return reserve_rings(ctx, n_steps);
}

int chiml_gpu_download_emitter_state(ChimlCtx* ctx, int slot, int sys, int which, double* out)
{
    if(!ctx || !out) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "download_emitter_state before commit");
    if(slot < 0 || slot >= (int)ctx->emitters.size()) return fail(ctx, CHIML_ERR_ARG, "download_emitter_state: bad slot");
    const EmitterDev& em = ctx->emitters[slot];
    if(sys < 0 || sys >= em.d.nsys || which < 0 || which > 4) return fail(ctx, CHIML_ERR_ARG, "download_emitter_state: bad sys/which");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const double* src = which == 0 ? em.d_rho : em.d_f[(em.fbase + which - 1) % 4];
    const size_t ne = (size_t)em.d.nemit, n2r = (size_t)em.n2 * 2;
    if(ne == 0) return CHIML_OK;
    if(em.group)      // AoS (sys, e, k, re/im): the layout asked for
    {
        CK(cudaMemcpy(out, src + (size_t)sys * n2r * ne, n2r * ne * sizeof(double), cudaMemcpyDeviceToHost));
        return CHIML_OK;
    }
    std::vector<double> soa(n2r * ne);
    CK(cudaMemcpy(soa.data(), src + (size_t)sys * n2r * ne, soa.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for(size_t e = 0; e < ne; ++e)
        for(size_t k = 0; k < n2r; ++k) out[e * n2r + k] = soa[k * ne + e];
    return CHIML_OK;
}

int chiml_gpu_download_emitter_pol(ChimlCtx* ctx, int slot, int comp, double* out)
{
    if(!ctx || !out) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "download_emitter_pol before commit");
    if(slot < 0 || slot >= (int)ctx->emitters.size() || comp < 0 || comp > 2) return fail(ctx, CHIML_ERR_ARG, "download_emitter_pol: bad slot/comp");
    const EmitterDev& em = ctx->emitters[slot];
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, em.d_P[comp], em.pbox * sizeof(double), cudaMemcpyDeviceToHost));
    return CHIML_OK;
}

int chiml_gpu_read_population(ChimlCtx* ctx, int slot, int det, double* out, size_t cap_samples, size_t* n_samples)
{
    if(!ctx) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "read_population before commit");
    if(slot < 0 || slot >= (int)ctx->emitters.size()) return fail(ctx, CHIML_ERR_ARG, "read_population: bad slot");
    const EmitterDev& em = ctx->emitters[slot];
    if(det < 0 || det >= em.d.npop) return fail(ctx, CHIML_ERR_ARG, "read_population: bad detector");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if(n_samples) *n_samples = em.pop_n - em.pop_base;
    const size_t n = std::min(cap_samples, em.pop_n - em.pop_base);
    for(size_t s = em.pop_base, done = 0; out && done < n; )
    {
        const size_t run = std::min(n - done, em.pop_cap - s % em.pop_cap);
        CK(cudaMemcpy(out + 2 * done, em.d_pop + ((size_t)det * em.pop_cap + s % em.pop_cap) * 2, run * 2 * sizeof(double), cudaMemcpyDeviceToHost));
        s += run; done += run;
    }
    return CHIML_OK;
}

int chiml_gpu_read_detector_range(ChimlCtx* ctx, int slot, size_t first, size_t n, double* out, size_t* n_read)
{
    if(!ctx || !out) return CHIML_ERR_ARG;
    if(!ctx->committed) return fail(ctx, CHIML_ERR_STATE, "read_detector_range before commit");
    if(slot < 0 || slot >= (int)ctx->detectors.size()) return fail(ctx, CHIML_ERR_ARG, "read_detector_range: bad slot");
    CK(cudaSetDevice(ctx->device));
    const DetectorDev& dt = ctx->detectors[slot];
    const size_t avail = (first >= dt.base && first < dt.count) ? dt.count - first : 0;    // consumed samples are gone
    const size_t m = std::min(n, avail);
    if(n_read) *n_read = m;
    CK(cudaStreamSynchronize(ctx->stream));
    return m ? ring_read(ctx, dt, first, m, out) : CHIML_OK;
}

} // extern "C"
