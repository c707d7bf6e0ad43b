// lbm_api.cu — the C ABI of include/neon_lbm.h: argument checking, layout, dispatch.  No CPU fallback anywhere: every
// compute entry point needs a CUDA device and fails with NLBM_ERR_CUDA otherwise.
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cuda.h>

#include "lbm_block.cuh"
#include "lbm_common.cuh"
#include "lbm_host.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
int cudaFail(cudaError_t e, const char* what)
{
    return fail(NLBM_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

int checkDesc(const nlbm_dense_desc* d, int elemBytes, bool needIn, bool needOut, bool needFlags)
{
    if (!d)
        return fail(NLBM_ERR_INVALID, "null descriptor");
    if (d->nx <= 0 || d->ny <= 0 || d->nz_local <= 0 || d->z_halo < 0 || d->z_halo > 1)
        return fail(NLBM_ERR_INVALID, "bad partition size %d x %d x %d (z_halo %d)", d->nx, d->ny, d->nz_local, d->z_halo);
    if (d->pitch_y < d->nx || d->pitch_y % 32 != 0 || d->pitch_y > INT_MAX)
        return fail(NLBM_ERR_INVALID, "pitch_y %lld must be a multiple of 32 elements and >= nx", (long long)d->pitch_y);
    if (d->pitch_z < d->pitch_y * d->ny || d->pitch_z % 32 != 0)
        return fail(NLBM_ERR_INVALID, "pitch_z %lld too small or misaligned", (long long)d->pitch_z);
    if (d->pitch_q < d->pitch_z * (d->nz_local + 2 * d->z_halo) || d->pitch_q % 32 != 0)
        return fail(NLBM_ERR_INVALID, "pitch_q %lld too small or misaligned", (long long)d->pitch_q);
    if (d->ny > 65535 || d->nz_local + 2 * d->z_halo > 65535)
        return fail(NLBM_ERR_UNSUPPORTED, "ny and nz_local are limited to 65535");
    if (needIn && (!d->pop_in || ((uintptr_t)d->pop_in & 127)))
        return fail(NLBM_ERR_INVALID, "pop_in null or not 128-byte aligned");
    if (needOut && (!d->pop_out || ((uintptr_t)d->pop_out & 127)))
        return fail(NLBM_ERR_INVALID, "pop_out null or not 128-byte aligned");
    if (needFlags && (!d->flags || ((uintptr_t)d->flags & 127)))
        return fail(NLBM_ERR_INVALID, "flags null or not 128-byte aligned");
    if (needIn && needOut && d->pop_in == d->pop_out)
        return fail(NLBM_ERR_INVALID, "pop_in and pop_out alias (the pull scheme needs two fields, LbmIteration.h:38-39)");
    (void)elemBytes;
    return NLBM_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encodeTiled()
{
    static EncodeTiledFn fn = nullptr;
    static bool          tried = false;
    if (!tried) {
        void*                           p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        tried = true;
    }
    return fn;
}

// 4-D map (x, y, memory plane, population) over a SoA population field; box = one tile of one population
CUtensorMapL2promotion l2Promotion(int sel)
{
    switch (sel) {
        case 1: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        case 2: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
        case 3: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
        default: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
}

// 3-D map (x, y, memory plane) over the flag words
bool makeFlagMap(CUtensorMap* m, const nlbm_dense_desc* d, int tx, int ty, int promo)
{
    EncodeTiledFn enc = encodeTiled();
    if (!enc)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)d->pitch_y, (cuuint64_t)d->ny, (cuuint64_t)(d->nz_local + 2 * d->z_halo)};
    const cuuint64_t strides[2] = {(cuuint64_t)d->pitch_y * 4, (cuuint64_t)d->pitch_z * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tx, (cuuint32_t)ty, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d->flags, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, l2Promotion(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool makePopMap(CUtensorMap* m, const nlbm_dense_desc* d, int q, int elemBytes, int tx, int ty, int promo)
{
    EncodeTiledFn enc = encodeTiled();
    if (!enc)
        return false;
    const cuuint64_t dims[4] = {(cuuint64_t)d->nx, (cuuint64_t)d->ny, (cuuint64_t)(d->nz_local + 2 * d->z_halo), (cuuint64_t)q};
    const cuuint64_t strides[3] = {(cuuint64_t)d->pitch_y * elemBytes, (cuuint64_t)d->pitch_z * elemBytes,
                                   (cuuint64_t)d->pitch_q * elemBytes};
    const cuuint32_t box[4] = {(cuuint32_t)tx, (cuuint32_t)ty, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, elemBytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d->pop_in, dims,
                     strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2Promotion(promo),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int smCount()
{
    static int sms[64] = {0};
    int        dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 0;
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return 0;
        sms[dev] = n;
    }
    return sms[dev];
}

int stepImpl(nlbm::StepKind kind, int elemBytes, const nlbm_dense_desc* d, double omega, int view, int opts, void* stream,
             const nlbm_peer_desc* peer = nullptr, int iterations = 0, const void* wallCacheIn = nullptr)
{
    if (int rc = checkDesc(d, elemBytes, true, true, true))
        return rc;
    if (view < NLBM_VIEW_STANDARD || view > NLBM_VIEW_BOUNDARY)
        return fail(NLBM_ERR_INVALID, "bad data view %d", view);
    const int arith = opts & 0xF;
    if (arith != NLBM_ARITH_REFERENCE && arith != NLBM_ARITH_FAST)
        return fail(NLBM_ERR_INVALID, "bad arithmetic mode %d", arith);
    nlbm::DenseArgs a;
    a.in = d->pop_in;
    a.out = d->pop_out;
    a.flags = d->flags;
    a.summary = nlbm::summaryPtr(*d);
    a.nx = d->nx;
    a.ny = d->ny;
    a.nzm = d->nz_local + 2 * d->z_halo;
    a.pitch_y = (int32_t)d->pitch_y;
    a.pitch_z = d->pitch_z;
    a.pitch_q = d->pitch_q;
    a.wpr = (int32_t)nlbm::summaryWordsPerRow(d->pitch_y);
    a.omega = omega;
    // how a thread learns about its cells: default the cell map; NLBM_OPT_FLAG_WORDS / NLBM_OPT_FLAGS_SUMMARY_FIRST select the others
    a.flagMode = ((opts >> 28) & 1) ? nlbm::kFlagWords : (((opts >> 20) & 1) ? nlbm::kFlagSummaryFirst : nlbm::kFlagCellMap);
    a.cellMap = nlbm::cellMapPtr(*d);
    a.prefetchXFaces = ((opts >> 27) & 1) ? 0 : 1;  // NLBM_OPT_NO_XFACE_PREFETCH
    a.specXFix = ((opts >> 29) & 1) ? 0 : 1;        // NLBM_OPT_NO_XFACE_FIXUP_PREFETCH
    a.keepCache = d->wall_cache;
    a.experiment = (opts >> 24) & 0x7;  // NLBM_OPT_EXPERIMENT: measurement only
    if (a.experiment)
        a.flagMode = nlbm::kFlagWords;
    a.lprLog2 = 5;
    a.peerMode = 0;
    a.nzLocal = d->nz_local;
    a.peer[0] = a.peer[1] = nullptr;
    a.peerOff[0] = a.peerOff[1] = a.peerPitchQ[0] = a.peerPitchQ[1] = 0;
    a.peerFlag[0] = a.peerFlag[1] = nullptr;
    a.counter = nullptr;
    a.signalValue = a.warpsPerFace = 0;
    if (peer) {
        if (view != NLBM_VIEW_STANDARD)
            return fail(NLBM_ERR_INVALID, "the fused step updates the whole partition");
        if (d->z_halo != 1 || d->nz_local < 2)
            return fail(NLBM_ERR_INVALID, "the fused step needs ghost planes and at least two local planes");
        if (!peer->counters || ((uintptr_t)peer->counters & 3))
            return fail(NLBM_ERR_INVALID, "counters null or misaligned");
        if ((peer->down_field && (!peer->down_flag || peer->down_nz_local < 1)) || (peer->up_field && (!peer->up_flag || peer->up_nz_local < 1)))
            return fail(NLBM_ERR_INVALID, "a neighbour needs its field, flag word and slab height");
        a.peerMode = 1;
        a.peer[0] = peer->down_field;
        a.peer[1] = peer->up_field;
        // my plane 0 lands in the lower neighbour's UPPER ghost plane (memory plane nz + 1), my top plane in the upper
        // neighbour's LOWER ghost plane (memory plane 0)
        a.peerOff[0] = (int64_t)(peer->down_nz_local + 1) * d->pitch_z;
        a.peerOff[1] = 0;
        a.peerPitchQ[0] = d->pitch_z * (peer->down_nz_local + 2);
        a.peerPitchQ[1] = d->pitch_z * (peer->up_nz_local + 2);
        a.peerFlag[0] = peer->down_flag;
        a.peerFlag[1] = peer->up_flag;
        a.counter = peer->counters;
        a.signalValue = peer->value;
    }
    // Views split at stencil radius 1 (both lattices): INTERNAL = local z in [1, nz-1), BOUNDARY = {0, nz-1}
    // (the reference's BOUNDARY span folds wrongly; this is the intended cover).
    const int r = 1, nz = d->nz_local;
    int       nzView = nz;
    a.zm0 = d->z_halo;
    a.fold = INT_MAX;
    a.skip = 0;
    if (view == NLBM_VIEW_INTERNAL) {
        a.zm0 = d->z_halo + r;
        nzView = nz - 2 * r;
    } else if (view == NLBM_VIEW_BOUNDARY) {
        if (nz >= 2 * r) {
            nzView = 2 * r;
            a.fold = r;
            a.skip = nz - 2 * r;
        }
    }
    if (nzView <= 0)
        return NLBM_OK;
    nlbm::StepLaunch l;
    l.nzView = nzView;
    l.vec = (opts >> 4) & 0xF;
    l.rowsLog2 = (opts >> 8) & 0xF;
    l.rpwSel = (opts >> 21) & 0x7;
    l.tmapA = nullptr;
    l.tmapB = nullptr;
    l.tmapF = nullptr;
    l.groups = (opts >> 18) & 0x3;
    l.numSms = 0;
    l.exact = ((opts >> 30) & 1) ? 0 : 1;  // NLBM_OPT_REF_LITERAL: the operand-for-operand transcription instead
    const int   kernelSel = peer ? 1 : (opts >> 12) & 0xF;  // 0 auto, 1 direct loads, 2 TMA-fed persistent
    const int   q = (kind == nlbm::kD3Q27_F32 || kind == nlbm::kD3Q27_F64) ? 27 : 19;
    CUtensorMap tmapA, tmapB, tmapF;
    const int   promo = (opts >> 16) & 0x3;
    // auto = the direct kernel: since all its loads are issued up front it beats the TMA-fed variant on every measured
    // configuration (profiles/); the TMA kernel stays selectable
    if (kernelSel == 2) {
        int tx, ty;
        nlbm::tmaTileShape(elemBytes, d->nx, &tx, &ty);
        const int sms = smCount();
        if (sms > 0 && makePopMap(&tmapA, d, q, elemBytes, tx, ty, promo) &&
            makePopMap(&tmapB, d, q, elemBytes, tx + 16 / elemBytes, ty, promo) && makeFlagMap(&tmapF, d, tx, ty, promo)) {
            l.tmapA = &tmapA;
            l.tmapB = &tmapB;
            l.tmapF = &tmapF;
            l.numSms = sms;
        } else if (kernelSel == 2) {
            return fail(NLBM_ERR_CUDA, "cannot build the TMA descriptor (cuTensorMapEncodeTiled unavailable or refused)");
        }
    } else if (kernelSel != 0 && kernelSel != 1) {
        return fail(NLBM_ERR_INVALID, "bad kernel selector %d", kernelSel);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (iterations > 0) {  // nlbm_dense_step_n: several iterations, the two fields swapping roles, without returning to the stream
        if (view != NLBM_VIEW_STANDARD || peer != nullptr || l.tmapA != nullptr)
            return fail(NLBM_ERR_INVALID, "multi-iteration launch: STANDARD view, direct kernel, no face push");
        nlbm::MultiArgs m{};
        m.fieldB = d->pop_out;
        m.keepCacheA = wallCacheIn;
        m.iterations = iterations;
        const int ce = (opts >> 16) & 0xF;  // NLBM_OPT_CHAIN_EARLY
        m.chainEarly = ce == 0 ? 0 : (ce == 15 ? -1 : 1 << (ce - 1));
        cudaError_t e = arith == NLBM_ARITH_REFERENCE ? nlbm::launchMultiRef(kind, a, m, l, st) : nlbm::launchMultiFast(kind, a, m, l, st);
        if (e != cudaSuccess)
            return cudaFail(e, "dense multi-iteration launch");
        return NLBM_OK;
    }
    cudaError_t  e = arith == NLBM_ARITH_REFERENCE ? nlbm::launchStepRef(kind, a, l, st) : nlbm::launchStepFast(kind, a, l, st);
    if (e != cudaSuccess)
        return cudaFail(e, "dense step launch");
    return NLBM_OK;
}

int crossing(int latticeQ, int ncomp, int dir, int* list)
{
    int n = 0;
    if (latticeQ == 0) {
        for (int q = 0; q < ncomp; ++q)
            list[n++] = q;
    } else if (latticeQ == 19) {
        for (int q = 0; q < 19; ++q)
            if (nlbm::Lattice<19>::c(q, 2) == dir)
                list[n++] = q;
    } else {
        for (int q = 0; q < 27; ++q)
            if (nlbm::Lattice<27>::c(q, 2) == dir)
                list[n++] = q;
    }
    return n;
}

int checkHalo(const nlbm_dense_desc* d, int elemBytes, int ncomp, int latticeQ, int dir)
{
    if (int rc = checkDesc(d, elemBytes, false, false, false))
        return rc;
    if (elemBytes != 4 && elemBytes != 8)
        return fail(NLBM_ERR_INVALID, "elem_bytes must be 4 or 8");
    if (dir != 1 && dir != -1)
        return fail(NLBM_ERR_INVALID, "dir must be +1 or -1");
    if (!(latticeQ == 0 || latticeQ == 19 || latticeQ == 27) || ncomp < 1 || ncomp > 27 || (latticeQ && ncomp != latticeQ))
        return fail(NLBM_ERR_INVALID, "bad component count %d / lattice %d", ncomp, latticeQ);
    return NLBM_OK;
}

}  // namespace

extern "C" {

int         nlbm_abi_version(void) { return NLBM_ABI_VERSION; }
const char* nlbm_last_error(void) { return g_err; }

int nlbm_device_count(void)
{
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaFail(e, "cudaGetDeviceCount");
        return -1;
    }
    return n;
}

int nlbm_dense_layout(nlbm_dense_desc* d, int q, int elem_bytes, size_t* pop_bytes, size_t* flag_bytes)
{
    if (!d || d->nx <= 0 || d->ny <= 0 || d->nz_local <= 0 || d->z_halo < 0 || d->z_halo > 1)
        return fail(NLBM_ERR_INVALID, "bad partition size");
    if ((elem_bytes != 4 && elem_bytes != 8) || q < 1 || q > 27)
        return fail(NLBM_ERR_INVALID, "bad element size %d or component count %d", elem_bytes, q);
    // rows start on 512-byte boundaries: one warp moves 32 x 16 bytes of a row per request
    d->pitch_y = nlbm::alignUp(d->nx, 512 / elem_bytes);
    d->pitch_z = d->pitch_y * d->ny;
    d->pitch_q = d->pitch_z * (d->nz_local + 2 * d->z_halo);
    if (pop_bytes)
        *pop_bytes = (size_t)q * (size_t)d->pitch_q * (size_t)elem_bytes;
    if (flag_bytes) {
        *flag_bytes = (size_t)nlbm::flagBufferBytes(*d);  // flag words + row summary + cell map
    }
    return NLBM_OK;
}

int nlbm_dense_wall_cache_layout(const nlbm_dense_desc* d, int q, int elem_bytes, size_t* bytes)
{
    if (!d || !bytes || d->ny <= 0 || d->nz_local <= 0 || d->z_halo < 0 || d->z_halo > 1 || q < 1 || q > 27 || (elem_bytes != 4 && elem_bytes != 8))
        return fail(NLBM_ERR_INVALID, "bad wall-cache request");
    *bytes = (size_t)nlbm::alignUp((int64_t)2 * q * (d->nz_local + 2 * d->z_halo) * d->ny * elem_bytes, 128);
    return NLBM_OK;
}
int nlbm_dense_wall_cache_build(const nlbm_dense_desc* d, int q, int elem_bytes, void* stream)
{
    if (int rc = checkDesc(d, elem_bytes, false, true, false))
        return rc;
    if (!d->wall_cache || ((uintptr_t)d->wall_cache & 127) || q < 1 || q > 27 || (elem_bytes != 4 && elem_bytes != 8))
        return fail(NLBM_ERR_INVALID, "wall cache null / misaligned, or bad component count / element size");
    cudaError_t e = nlbm::launchWallCacheBuild(*d, q, elem_bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "wall cache build launch");
}

int nlbm_dense_classify(const nlbm_dense_desc* d, int geom, const double* sphere, void* stream)
{
    if (int rc = checkDesc(d, 4, false, false, true))
        return rc;
    if (geom < 0 || geom > 2)
        return fail(NLBM_ERR_INVALID, "bad geometry %d", geom);
    if (d->nx != d->gnx || d->ny != d->gny)
        return fail(NLBM_ERR_INVALID, "partitions split z only (dGrid_imp.h:32-63): nx,ny must equal gnx,gny");
    cudaError_t e = nlbm::launchClassify(*d, geom, sphere, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "classify launch");
}

int nlbm_dense_flags_commit(const nlbm_dense_desc* d, void* stream)
{
    if (int rc = checkDesc(d, 4, false, false, true))
        return rc;
    cudaError_t e = nlbm::launchSummary(*d, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "flag summary launch");
}

int nlbm_selftest_exact(int kind, uint64_t n, uint64_t seed, uint64_t* mismatches)
{
    if ((kind != 0 && kind != 1) || mismatches == nullptr)
        return fail(NLBM_ERR_INVALID, "selftest kind %d", kind);
    unsigned long long* d = nullptr;
    cudaError_t         e = cudaMalloc(reinterpret_cast<void**>(&d), sizeof(unsigned long long));
    if (e != cudaSuccess)
        return cudaFail(e, "selftest alloc");
    e = cudaMemset(d, 0, sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = nlbm::launchSelftestExact(kind, n, seed, d, nullptr);
    unsigned long long h = 0;
    if (e == cudaSuccess)
        e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
        return cudaFail(e, "selftest");
    *mismatches = h;
    return NLBM_OK;
}

int nlbm_dense_flags_from_classes(const nlbm_dense_desc* d, const uint8_t* classes, int zm_first, int nplanes, void* stream)
{
    if (int rc = checkDesc(d, 4, false, false, true))
        return rc;
    if (classes == nullptr || nplanes < 0 || zm_first < 0 || zm_first + nplanes > d->nz_local + 2 * d->z_halo)
        return fail(NLBM_ERR_INVALID, "classes cover memory planes [%d, %d) of %d", zm_first, zm_first + nplanes, d->nz_local + 2 * d->z_halo);
    cudaError_t e = nlbm::launchFlagsFromClasses(*d, classes, zm_first, nplanes, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "flags-from-classes launch");
}

int nlbm_dense_wall_mask(const nlbm_dense_desc* d, int q, int32_t* d_bad, void* stream)
{
    if (int rc = checkDesc(d, 4, false, false, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchWallMask(*d, q, d_bad, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "wall mask launch");
}

int nlbm_dense_init_pop_f32(const nlbm_dense_desc* d, int q, double ulb, void* stream)
{
    if (int rc = checkDesc(d, 4, false, true, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchInitPop<float>(*d, q, ulb, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "init launch");
}
int nlbm_dense_init_pop_f64(const nlbm_dense_desc* d, int q, double ulb, void* stream)
{
    if (int rc = checkDesc(d, 8, false, true, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchInitPop<double>(*d, q, ulb, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "init launch");
}

int nlbm_d3q19_f32_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream)
{
    return stepImpl(nlbm::kD3Q19_F32, 4, d, omega, data_view, opts, stream);
}
int nlbm_d3q19_f64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream)
{
    return stepImpl(nlbm::kD3Q19_F64, 8, d, omega, data_view, opts, stream);
}
int nlbm_d3q19_f32c64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream)
{
    return stepImpl(nlbm::kD3Q19_F32C64, 4, d, omega, data_view, opts, stream);
}
int nlbm_d3q27_f32_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream)
{
    return stepImpl(nlbm::kD3Q27_F32, 4, d, omega, data_view, opts, stream);
}
int nlbm_d3q27_f64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream)
{
    return stepImpl(nlbm::kD3Q27_F64, 8, d, omega, data_view, opts, stream);
}

int nlbm_dense_step_n(int kind, const nlbm_dense_desc* d, const void* wall_cache_in, double omega, int iterations, int opts, void* stream)
{
    if (kind < 0 || kind > 4)
        return fail(NLBM_ERR_INVALID, "bad step kind %d", kind);
    if (iterations < 1)
        return fail(NLBM_ERR_INVALID, "iterations must be >= 1");
    if (!d || d->z_halo != 0)
        return fail(NLBM_ERR_UNSUPPORTED, "multi-iteration launch needs a partition without neighbours (z_halo = 0): nothing is exchanged between its iterations");
    const int elem = (kind == nlbm::kD3Q19_F64 || kind == nlbm::kD3Q27_F64) ? 8 : 4;
    return stepImpl((nlbm::StepKind)kind, elem, d, omega, NLBM_VIEW_STANDARD, opts, stream, nullptr, iterations, wall_cache_in);
}

int nlbm_dense_step_push(int kind, const nlbm_dense_desc* d, const nlbm_peer_desc* peer, double omega, int opts, void* stream)
{
    if (kind < 0 || kind > 4)
        return fail(NLBM_ERR_INVALID, "bad kind %d", kind);
    if (!peer)
        return fail(NLBM_ERR_INVALID, "null peer descriptor");
    const int elem = (kind == nlbm::kD3Q19_F64 || kind == nlbm::kD3Q27_F64) ? 8 : 4;
    return stepImpl((nlbm::StepKind)kind, elem, d, omega, NLBM_VIEW_STANDARD, opts, stream, peer);
}

int nlbm_d3q19_f32_dense_rho_u(const nlbm_dense_desc* d, void* rho, void* u, void* stream)
{
    if (int rc = checkDesc(d, 4, true, false, true))
        return rc;
    if (!rho || !u)
        return fail(NLBM_ERR_INVALID, "null output");
    cudaError_t e = nlbm::launchRhoU<float, float>(*d, rho, u, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "rho/u launch");
}
int nlbm_d3q19_f64_dense_rho_u(const nlbm_dense_desc* d, void* rho, void* u, void* stream)
{
    if (int rc = checkDesc(d, 8, true, false, true))
        return rc;
    if (!rho || !u)
        return fail(NLBM_ERR_INVALID, "null output");
    cudaError_t e = nlbm::launchRhoU<double, double>(*d, rho, u, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "rho/u launch");
}

int nlbm_dense_halo_push(const nlbm_dense_desc* s, const void* src_field, const nlbm_dense_desc* t, void* dst_field,
                         int elem_bytes, int ncomp, int lattice_q, int dir, void* stream)
{
    if (int rc = checkHalo(s, elem_bytes, ncomp, lattice_q, dir))
        return rc;
    if (int rc = checkHalo(t, elem_bytes, ncomp, lattice_q, dir))
        return rc;
    if (!src_field || !dst_field)
        return fail(NLBM_ERR_INVALID, "null field");
    if (t->z_halo < 1)
        return fail(NLBM_ERR_INVALID, "destination partition has no ghost planes");
    if (s->pitch_y != t->pitch_y || s->pitch_z != t->pitch_z || s->ny != t->ny || s->nx != t->nx)
        return fail(NLBM_ERR_INVALID, "source and destination planes differ in shape");
    int             list[27];
    nlbm::PlaneList pl;
    pl.n = crossing(lattice_q, ncomp, dir, list);
    const int64_t zs = dir > 0 ? s->z_halo + s->nz_local - 1 : s->z_halo;
    const int64_t zd = dir > 0 ? 0 : t->z_halo + t->nz_local;
    for (int i = 0; i < pl.n; ++i) {
        pl.src[i] = (list[i] * s->pitch_q + zs * s->pitch_z) * elem_bytes;
        pl.dst[i] = (list[i] * t->pitch_q + zd * t->pitch_z) * elem_bytes;
    }
    cudaError_t e = nlbm::launchPlaneCopy(src_field, dst_field, pl, (size_t)s->pitch_z * elem_bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "halo push launch");
}

int nlbm_dense_halo_push2(const nlbm_dense_desc* s, const void* src_field, void* up_field, int32_t up_nz_local, uint32_t* up_flag,
                          void* down_field, int32_t down_nz_local, uint32_t* down_flag, uint32_t* counter, uint32_t value,
                          int elem_bytes, int ncomp, int lattice_q, void* stream)
{
    if (int rc = checkHalo(s, elem_bytes, ncomp, lattice_q, 1))
        return rc;
    if (!src_field)
        return fail(NLBM_ERR_INVALID, "null field");
    if (s->z_halo < 1)
        return fail(NLBM_ERR_INVALID, "the partition has no z-neighbours (z_halo = 0)");
    if (!counter || ((uintptr_t)counter & 3))
        return fail(NLBM_ERR_INVALID, "counter null or misaligned");
    if ((up_field && (!up_flag || up_nz_local < 1)) || (down_field && (!down_flag || down_nz_local < 1)))
        return fail(NLBM_ERR_INVALID, "a neighbour needs its field, flag word and slab height");
    if (((uintptr_t)up_flag | (uintptr_t)down_flag) & 3)
        return fail(NLBM_ERR_INVALID, "flag pointer misaligned");
    int             list[27];
    nlbm::PlaneList pl;
    pl.n = 0;
    int nUp = 0;
    if (up_field) {  // my top plane -> the lower ghost plane (memory plane 0) of the neighbour above
        const int     n = crossing(lattice_q, ncomp, +1, list);
        const int64_t zs = s->z_halo + s->nz_local - 1, pq = s->pitch_z * (up_nz_local + 2);
        for (int i = 0; i < n; ++i, ++pl.n) {
            pl.src[pl.n] = (list[i] * s->pitch_q + zs * s->pitch_z) * elem_bytes;
            pl.dst[pl.n] = (list[i] * pq) * elem_bytes;
        }
        nUp = n;
    }
    if (down_field) {  // my bottom plane -> the upper ghost plane (memory plane nz + 1) of the neighbour below
        const int     n = crossing(lattice_q, ncomp, -1, list);
        const int64_t zs = s->z_halo, pq = s->pitch_z * (down_nz_local + 2), zd = down_nz_local + 1;
        for (int i = 0; i < n; ++i, ++pl.n) {
            pl.src[pl.n] = (list[i] * s->pitch_q + zs * s->pitch_z) * elem_bytes;
            pl.dst[pl.n] = (list[i] * pq + zd * s->pitch_z) * elem_bytes;
        }
    }
    cudaError_t e = nlbm::launchFacePush2(src_field, pl, nUp, up_field, down_field, up_field ? up_flag : nullptr,
                                          down_field ? down_flag : nullptr, counter, value, (size_t)s->pitch_z * elem_bytes,
                                          (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "halo push2 launch");
}

int nlbm_dense_halo_pack(const nlbm_dense_desc* d, const void* field, int elem_bytes, int ncomp, int lattice_q, int dir,
                         void* buffer, size_t* bytes, void* stream)
{
    if (int rc = checkHalo(d, elem_bytes, ncomp, lattice_q, dir))
        return rc;
    int             list[27];
    nlbm::PlaneList pl;
    pl.n = crossing(lattice_q, ncomp, dir, list);
    const size_t planeBytes = (size_t)d->pitch_z * elem_bytes;
    if (bytes)
        *bytes = planeBytes * pl.n;
    if (!buffer)
        return NLBM_OK;  // size query
    if (!field)
        return fail(NLBM_ERR_INVALID, "null field");
    const int64_t zs = dir > 0 ? d->z_halo + d->nz_local - 1 : d->z_halo;
    for (int i = 0; i < pl.n; ++i) {
        pl.src[i] = (list[i] * d->pitch_q + zs * d->pitch_z) * elem_bytes;
        pl.dst[i] = (int64_t)i * planeBytes;
    }
    cudaError_t e = nlbm::launchPlaneCopy(field, buffer, pl, planeBytes, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "halo pack launch");
}

int nlbm_dense_halo_unpack(const nlbm_dense_desc* d, void* field, int elem_bytes, int ncomp, int lattice_q, int dir,
                           const void* buffer, void* stream)
{
    if (int rc = checkHalo(d, elem_bytes, ncomp, lattice_q, dir))
        return rc;
    if (!field || !buffer)
        return fail(NLBM_ERR_INVALID, "null field or buffer");
    if (d->z_halo < 1)
        return fail(NLBM_ERR_INVALID, "partition has no ghost planes");
    int             list[27];
    nlbm::PlaneList pl;
    pl.n = crossing(lattice_q, ncomp, dir, list);
    const size_t  planeBytes = (size_t)d->pitch_z * elem_bytes;
    const int64_t zd = dir > 0 ? 0 : d->z_halo + d->nz_local;  // data moving up lands in my lower ghost plane
    for (int i = 0; i < pl.n; ++i) {
        pl.src[i] = (int64_t)i * planeBytes;
        pl.dst[i] = (list[i] * d->pitch_q + zd * d->pitch_z) * elem_bytes;
    }
    cudaError_t e = nlbm::launchPlaneCopy(buffer, field, pl, planeBytes, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "halo unpack launch");
}

// ------------------------------------------------------------------------------------------------ block-sparse path
static int checkBlockDesc(const nlbm_block_desc* d, bool needIn, bool needOut, bool needFlags)
{
    if (!d)
        return fail(NLBM_ERR_INVALID, "null descriptor");
    if (d->n_blocks > d->n_blocks_alloc || d->n_blocks_alloc == 0xFFFFFFFFu)
        return fail(NLBM_ERR_INVALID, "bad block counts %u local of %u allocated", d->n_blocks, d->n_blocks_alloc);
    if ((uint64_t)d->n_down + d->n_up > d->n_blocks)
        return fail(NLBM_ERR_INVALID, "boundary layers (%u + %u blocks) exceed the %u local blocks", d->n_down, d->n_up, d->n_blocks);
    if (!d->info || ((uintptr_t)d->info & 127))
        return fail(NLBM_ERR_INVALID, "info null or not 128-byte aligned");
    if (needIn && (!d->pop_in || ((uintptr_t)d->pop_in & 127)))
        return fail(NLBM_ERR_INVALID, "pop_in null or not 128-byte aligned");
    if (needOut && (!d->pop_out || ((uintptr_t)d->pop_out & 127)))
        return fail(NLBM_ERR_INVALID, "pop_out null or not 128-byte aligned");
    if (needFlags && (!d->flags || ((uintptr_t)d->flags & 127)))
        return fail(NLBM_ERR_INVALID, "flags null or not 128-byte aligned");
    if (needIn && needOut && d->pop_in == d->pop_out)
        return fail(NLBM_ERR_INVALID, "pop_in and pop_out alias (the pull scheme needs two fields, LbmIteration.h:38-39)");
    return NLBM_OK;
}

static int blockStepImpl(nlbm::StepKind kind, const nlbm_block_desc* d, double omega, int view, int opts, void* stream)
{
    if (int rc = checkBlockDesc(d, true, true, true))
        return rc;
    if (view < NLBM_VIEW_STANDARD || view > NLBM_VIEW_BOUNDARY)
        return fail(NLBM_ERR_INVALID, "bad data view %d", view);
    const int arith = opts & 0xF;
    if (arith != NLBM_ARITH_REFERENCE && arith != NLBM_ARITH_FAST)
        return fail(NLBM_ERR_INVALID, "bad arithmetic mode %d", arith);
    nlbm::BlockArgs a;
    a.in = d->pop_in;
    a.out = d->pop_out;
    a.flags = d->flags;
    a.info = d->info;
    a.popPitch = (int64_t)d->n_blocks_alloc * nlbm::kBlockCells;
    a.eagerFlags = (opts >> 28) & 1;  // NLBM_OPT_FLAG_WORDS
    a.omega = omega;
    cudaStream_t st = (cudaStream_t)stream;
    // block ranges of the view: [0, n_down) lowest layer, [n_blocks - n_up, n_blocks) highest layer
    uint32_t ranges[2][2] = {{0, d->n_blocks}, {0, 0}};
    if (view == NLBM_VIEW_INTERNAL) {
        ranges[0][0] = d->n_down;
        ranges[0][1] = d->n_blocks - d->n_down - d->n_up;
    } else if (view == NLBM_VIEW_BOUNDARY) {
        ranges[0][1] = d->n_down;
        ranges[1][0] = d->n_blocks - d->n_up;
        ranges[1][1] = d->n_up;
    }
    for (auto& r : ranges) {
        if (r[1] == 0)
            continue;
        a.firstBlock = r[0];
        a.nBlocks = r[1];
        cudaError_t e = arith == NLBM_ARITH_REFERENCE ? nlbm::launchBlockStepRef(kind, a, r[1], st, ((opts >> 30) & 1) == 0)
                                                      : nlbm::launchBlockStepFast(kind, a, r[1], st);
        if (e != cudaSuccess)
            return cudaFail(e, "block step launch");
    }
    return NLBM_OK;
}

int nlbm_d3q19_f32_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream)
{
    return blockStepImpl(nlbm::kD3Q19_F32, d, omega, data_view, opts, stream);
}
int nlbm_d3q19_f64_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream)
{
    return blockStepImpl(nlbm::kD3Q19_F64, d, omega, data_view, opts, stream);
}
int nlbm_d3q27_f32_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream)
{
    return blockStepImpl(nlbm::kD3Q27_F32, d, omega, data_view, opts, stream);
}
int nlbm_d3q27_f64_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream)
{
    return blockStepImpl(nlbm::kD3Q27_F64, d, omega, data_view, opts, stream);
}

int nlbm_block_classify(const nlbm_block_desc* d, int geom, const double* sphere, const uint32_t* active_mask, void* stream)
{
    if (int rc = checkBlockDesc(d, false, false, true))
        return rc;
    if (geom < 0 || geom > 2)
        return fail(NLBM_ERR_INVALID, "bad geometry %d", geom);
    cudaError_t e = nlbm::launchBlockClassify(*d, geom, sphere, active_mask, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "block classify launch");
}

int nlbm_block_wall_mask(const nlbm_block_desc* d, int q, int32_t* d_bad, void* stream)
{
    if (int rc = checkBlockDesc(d, false, false, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchBlockWallMask(*d, q, d_bad, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "block wall mask launch");
}

int nlbm_block_init_pop_f32(const nlbm_block_desc* d, int q, double ulb, void* stream)
{
    if (int rc = checkBlockDesc(d, false, true, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchBlockInitPop<float>(*d, q, ulb, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "block init launch");
}
int nlbm_block_init_pop_f64(const nlbm_block_desc* d, int q, double ulb, void* stream)
{
    if (int rc = checkBlockDesc(d, false, true, true))
        return rc;
    if (q != 19 && q != 27)
        return fail(NLBM_ERR_INVALID, "lattice must be 19 or 27");
    cudaError_t e = nlbm::launchBlockInitPop<double>(*d, q, ulb, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "block init launch");
}

int nlbm_block_halo_push(const nlbm_block_desc* s, const void* src_field, const nlbm_block_desc* t, void* dst_field,
                         uint32_t dst_first_ghost, int elem_bytes, int ncomp, int lattice_q, int dir, void* stream)
{
    if (int rc = checkBlockDesc(s, false, false, false))
        return rc;
    if (int rc = checkBlockDesc(t, false, false, false))
        return rc;
    if (!src_field || !dst_field)
        return fail(NLBM_ERR_INVALID, "null field");
    if (elem_bytes != 4 && elem_bytes != 8)
        return fail(NLBM_ERR_INVALID, "elem_bytes must be 4 or 8");
    if (dir != 1 && dir != -1)
        return fail(NLBM_ERR_INVALID, "dir must be +1 or -1");
    if (!(lattice_q == 0 || lattice_q == 19 || lattice_q == 27) || ncomp < 1 || ncomp > 27 || (lattice_q && ncomp != lattice_q))
        return fail(NLBM_ERR_INVALID, "bad component count %d / lattice %d", ncomp, lattice_q);
    const uint32_t n = dir > 0 ? s->n_up : s->n_down;
    if ((uint64_t)dst_first_ghost + n > t->n_blocks_alloc || dst_first_ghost < t->n_blocks)
        return fail(NLBM_ERR_INVALID, "ghost range [%u, %u) outside the destination's ghost blocks", dst_first_ghost, dst_first_ghost + n);
    int       list[27];
    const int nc = crossing(lattice_q, ncomp, dir, list);
    const uint32_t first = dir > 0 ? s->n_blocks - s->n_up : 0;
    cudaError_t    e = nlbm::launchBlockSliceCopy(src_field, dst_field, elem_bytes, list, nc, (int64_t)s->n_blocks_alloc * 512,
                                                  (int64_t)t->n_blocks_alloc * 512, first, dst_first_ghost, n, dir > 0 ? 7 : 0,
                                                  (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "block halo push launch");
}

int nlbm_ipc_export(const void* ptr, unsigned char* handle64, uint64_t* offset)
{
    if (!ptr || !handle64 || !offset)
        return fail(NLBM_ERR_INVALID, "null argument");
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range = nullptr;
    if (!range) {
        void*                           p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return fail(NLBM_ERR_CUDA, "cuMemGetAddressRange unavailable");
        range = (RangeFn)p;
    }
    CUdeviceptr base = 0;
    size_t      size = 0;
    if (range(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS)
        return fail(NLBM_ERR_CUDA, "cuMemGetAddressRange failed: not a device allocation");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t        e = cudaIpcGetMemHandle(&h, (void*)base);
    if (e != cudaSuccess)
        return cudaFail(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, 64);
    *offset = (uint64_t)((CUdeviceptr)ptr - base);
    return NLBM_OK;
}

int nlbm_ipc_import(const unsigned char* handle64, void** base)
{
    if (!handle64 || !base)
        return fail(NLBM_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "cudaIpcOpenMemHandle");
}

int nlbm_ipc_close(void* base)
{
    cudaError_t e = cudaIpcCloseMemHandle(base);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "cudaIpcCloseMemHandle");
}

int nlbm_enable_peer_access(int peer_device)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return cudaFail(e, "cudaGetDevice");
    if (peer_device == dev)
        return NLBM_OK;
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
    if (e != cudaSuccess)
        return cudaFail(e, "cudaDeviceCanAccessPeer");
    if (!can)
        return fail(NLBM_ERR_UNSUPPORTED, "device %d cannot access device %d directly", dev, peer_device);
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();  // clear the sticky-less error state
        return NLBM_OK;
    }
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "cudaDeviceEnablePeerAccess");
}

int nlbm_flag_signal(uint32_t* flag, uint32_t value, void* stream)
{
    if (!flag || ((uintptr_t)flag & 3))
        return fail(NLBM_ERR_INVALID, "flag pointer null or misaligned");
    cudaError_t e = nlbm::launchFlagSignal(flag, value, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "flag signal launch");
}

int nlbm_flag_wait(const uint32_t* flag, uint32_t value, uint32_t timeout_ms, int32_t* d_err, void* stream)
{
    if (!flag || ((uintptr_t)flag & 3))
        return fail(NLBM_ERR_INVALID, "flag pointer null or misaligned");
    if (timeout_ms == 0)
        return fail(NLBM_ERR_INVALID, "a wait needs a time-out");
    cudaError_t e = nlbm::launchFlagWait(flag, nullptr, value, timeout_ms, d_err, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "flag wait launch");
}

int nlbm_flag_wait2(const uint32_t* flag_a, const uint32_t* flag_b, uint32_t value, uint32_t timeout_ms, int32_t* d_err, void* stream)
{
    if ((!flag_a && !flag_b) || (((uintptr_t)flag_a | (uintptr_t)flag_b) & 3))
        return fail(NLBM_ERR_INVALID, "flag pointers null or misaligned");
    if (timeout_ms == 0)
        return fail(NLBM_ERR_INVALID, "a wait needs a time-out");
    cudaError_t e = nlbm::launchFlagWait(flag_a, flag_b, value, timeout_ms, d_err, (cudaStream_t)stream);
    return e == cudaSuccess ? NLBM_OK : cudaFail(e, "flag wait launch");
}

}  // extern "C"
