/* neon_lbm.h — C ABI of the B200-native (sm_100a) Neon LBM hot path.
 *
 * This is the drop-in boundary.  In the reference the path is
 * entered through
 *     Neon::set::Container::run(int streamIdx, Neon::DataView)
 *         libNeonSet/include/Neon/set/Containter.h:25-27
 *     -> DeviceContainer<Grid,Lambda>::run      container/DeviceContainer.h:88-111
 *     -> DevSet::launchLambdaOnSpan             libNeonSet/include/Neon/set/DevSet.h:226-261,356-399
 *     -> GpuDevice::kernel.cudaLaunchKernel     libNeonSys/include/Neon/sys/devices/gpu/GpuDevice.h:151-191
 * with the LBM lambda of benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:285-325
 * as payload, and the halo side through dField::newHaloUpdate
 * (libNeonDomain/include/Neon/domain/details/dGrid/dField.h:84-87).  Every entry
 * point below names the reference interface it replaces.
 *
 * Conventions (all entry points):
 *  - plain C, no C++/torch types; device buffers are BORROWED from the caller's
 *    Field (the reference's Field owns its MemSet, dField.h:150-195); nothing is
 *    allocated or freed inside step calls;
 *  - asynchronous: work is enqueued on `stream` (a cudaStream_t passed as void*)
 *    of the CURRENT device and the call returns; completion is observed through
 *    the caller's stream/event API (Backend::sync, Backend.h:230-261);
 *  - thread-safe, no global mutable state except the per-thread last-error
 *    string (one host thread per GPU, DevSet.h:372-391);
 *  - returns 0 on success, a non-zero nlbm_status otherwise and never throws;
 *    the C++ veneer converts non-zero into NeonException (GpuDevice.h:165-188).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with NLBM_ERR_CUDA.
 */
#ifndef NEON_LBM_H
#define NEON_LBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLBM_ABI_VERSION 3

typedef enum nlbm_status {
    NLBM_OK = 0,
    NLBM_ERR_INVALID = 1,  /* bad descriptor / argument                            */
    NLBM_ERR_CUDA = 2,     /* a CUDA runtime call failed (see nlbm_last_error)     */
    NLBM_ERR_GEOMETRY = 3, /* a bulk cell has a neighbour outside the domain       */
    NLBM_ERR_UNSUPPORTED = 4
} nlbm_status;

/* Neon::DataView, libNeonCore/include/Neon/core/types/DataView.h:7-12.
 * INTERNAL = local z in [r, nz_local-r), BOUNDARY = [0,r) U [nz_local-r, nz_local)
 * (r = z_halo; the reference's BOUNDARY span folds wrongly). */
typedef enum nlbm_data_view { NLBM_VIEW_STANDARD = 0, NLBM_VIEW_INTERNAL = 1, NLBM_VIEW_BOUNDARY = 2 } nlbm_data_view;

/* Cell classes — values of CellType::Classification,
 * benchmarks/lbm-lid-driven-cavity-flow/src/CellType.h:5-11.                    */
typedef enum nlbm_cell_class { NLBM_BOUNCE_BACK = 0, NLBM_MOVING_WALL = 1, NLBM_BULK = 2, NLBM_UNDEFINED = 3 } nlbm_cell_class;

/* Per-cell flag word (replaces the 8-byte CellType struct, CellType.h:33-34):
 *   bits  0..26  wallNghBitflag (bit k set <=> cell at x - c_k is not bulk)
 *   bits 28..29  classification                                                  */
#define NLBM_FLAG_MASK_BITS 0x07FFFFFFu
#define NLBM_FLAG_CLASS_SHIFT 28
#define NLBM_FLAG_CLASS(f) (((f) >> NLBM_FLAG_CLASS_SHIFT) & 3u)
/* bGrid only, bit 30 of the flag word of a block's FIRST cell: every cell of the block is bulk without a wall neighbour.  Set by
 * nlbm_block_wall_mask, cleared by whoever rewrites the word; the block step kernels then skip the block's flag words.  Not part
 * of the cell's flags: decode with NLBM_FLAG_CLASS / NLBM_FLAG_MASK_BITS.                                                  */
#define NLBM_FLAG_BLOCK_PLAIN 0x40000000u

/* Arithmetic mode (bits 0..3 of `opts`).
 * REFERENCE reproduces the rounding of the reference expressions bit for bit
 * (double literals promote fp32 expressions to double, LbmTools.h:216-254; no FMA
 * contraction; IEEE division).  FAST uses fused multiply-add in the storage
 * precision; results then agree with the reference within the north-star
 * tolerance (1e-5 relative fp32, 1e-12 fp64), not bitwise.                       */
#define NLBM_ARITH_REFERENCE 0
#define NLBM_ARITH_FAST 1
/* REFERENCE arithmetic for D3Q19 fp32 is evaluated by default through a re-arrangement that performs every rounding of the
 * reference expressions on the same exact value (same bits, proven per merge in csrc/lbm_collide_exact.cuh) with a third
 * of the float<->double conversions; this bit selects the operand-for-operand transcription instead.  Never changes
 * results.                                                                       */
#define NLBM_OPT_REF_LITERAL (1 << 30)
/* Tuning (bits 4..7: cells per thread along x, 0 = library default; bits 8..11:
 * log2 of rows per block, 0 = default).  Never changes results.                  */
#define NLBM_OPT_VEC(v) (((v)&0xF) << 4)
#define NLBM_OPT_ROWS_LOG2(r) (((r)&0xF) << 8)
/* bits 12..15: kernel — 0 library default, 1 direct (aligned 16-byte loads + warp shuffles for the x shift),
 * 2 persistent TMA-fed (cp.async.bulk.tensor tiles staged in shared memory, mbarrier ring).  Default: direct.   */
#define NLBM_OPT_KERNEL(k) (((k)&0xF) << 12)
/* experiment knobs of the TMA kernel (never change results): bits 16..17 L2 promotion of the tensor maps (0 256 B,
 * 1 128 B, 2 64 B, 3 none), bits 18..19 consumer groups per CTA (0 = default 3).                            */
#define NLBM_OPT_TMA_L2PROMO(p) (((p)&0x3) << 16)
#define NLBM_OPT_TMA_GROUPS(g) (((g)&0x3) << 18)
/* the same bits 16..19 in nlbm_dense_step_n (launch chain; never changes results): how many z planes of an iteration start on the
 * per-plane counters instead of waiting for the whole previous launch — 0 library default (one chip-load of blocks), 1..14:
 * 2^(e-1) planes, 15: every plane.                                                                              */
#define NLBM_OPT_CHAIN_EARLY(e) (((e)&0xF) << 16)
/* bit 20 (direct kernel): consult the row summary first and fetch flag words only where a 32-cell chunk holds a non-plain
 * cell (saves up to 4 B/cell of traffic).  Default (bit clear): the flag words travel with the populations — one dependent
 * memory round trip less for warps that touch walls, measured faster on B200.                                          */
#define NLBM_OPT_FLAGS_SUMMARY_FIRST (1 << 20)
/* bits 21..23 (direct kernel): rows a warp covers, 0 = library default, else log2(rows) + 1 (1: whole-row warps of 32 lanes,
 * 3: 4 rows x 8 lanes ...).  Narrow warp tiles keep the wall round trip away from most warps.  Never changes results.      */
#define NLBM_OPT_ROWS_PER_WARP_LOG2P1(r) (((r)&0x7) << 21)
/* bits 24..26 (direct kernel): MEASUREMENT ONLY — the results are WRONG.  Value 3: flags are honoured (who is updated) but wall
 * fix-ups and kept wall values are skipped — what the wall path costs.  (Values 1 "no flag loads, every cell plain bulk" and
 * 2 "flags loaded but ignored" were used for the attribution of profiles/r01s and are not wired in the shipped kernel.)
 * Used by bench.py --experiment; never set by the host layers.                                                           */
#define NLBM_OPT_EXPERIMENT(e) (((e)&0x7) << 24)
/* bit 27 (direct kernel): do not fetch the output-field values of the cells on the x faces of the box speculatively (default:
 * fetch them with the streaming loads — the kept wall values of those cells are then no dependent DRAM round trip).  Never
 * changes results.                                                                                                      */
#define NLBM_OPT_NO_XFACE_PREFETCH (1 << 27)
/* bit 28 (direct kernel): fetch the 4-byte flag word of EVERY cell together with the populations (the round-1 default).
 * Default (bit clear): each thread reads one byte of the cell map (1 byte per 4 cells, kept behind the flag words by the
 * set-up calls) with its populations and fetches flag words only where a bulk cell has wall bits or shares the thread with a
 * non-bulk cell — 0.25 instead of 4 B/cell of flag traffic.  Never changes results.
 * Block kernels: fetch every flag word with the populations and ignore NLBM_FLAG_BLOCK_PLAIN (default: blocks that carry it
 * load no flag words).                                                                                                  */
#define NLBM_OPT_FLAG_WORDS (1 << 28)
/* bit 29 (direct kernel): do not fetch the wall fix-up operands of the cells next to the x faces speculatively (default: they
 * travel with the streaming loads, and are used only if the cell's wall bits are exactly the x-face set).  Never changes results. */
#define NLBM_OPT_NO_XFACE_FIXUP_PREFETCH (1 << 29)
#define NLBM_KERNEL_AUTO 0
#define NLBM_KERNEL_DIRECT 1
#define NLBM_KERNEL_TMA 2

/* Dense (dGrid) partition descriptor: one z-slab of the global box on one GPU.
 * Replaces what the reference kernel receives by value: dSpan {dataView, zHalo,
 * zBoundary, dim} (dSpan.h:45-48) and dPartition members (dPartition.h:423-434).
 *
 * Memory layout (structure of arrays, one plane set per population):
 *   pop[q][zm][y][x] at  q*pitch_q + zm*pitch_z + y*pitch_y + x      (elements)
 *   zm = z_local + z_halo in [0, nz_local + 2*z_halo)   ghost planes at both ends
 *   pitch_y*sizeof(T) is a multiple of 128 bytes (nlbm_dense_layout picks 512: one warp
 *   request of 32 x 16 bytes); base pointers 128-byte aligned.
 * The flag array uses the same (zm,y,x) indexing with pitch_y / pitch_z.  Behind the
 * per-cell words the SAME buffer holds a small per-row summary (which 32-cell chunks
 * contain anything but plain bulk cells) and a cell map (one byte per 4 cells: which are
 * bulk, which are anything but plain bulk) that let the step kernels skip flag loads;
 * nlbm_dense_classify / nlbm_dense_wall_mask keep both current, and a caller that writes
 * flag words itself must call nlbm_dense_flags_commit before stepping.
 * (Reference: unpadded SoA, dField_imp.h:67-87.)                                 */
typedef struct nlbm_dense_desc {
    void*     pop_in;   /* device, borrowed: populations read  (fIn,  const STENCIL)  */
    void*     pop_out;  /* device, borrowed: populations written (fOut, MAP)          */
    uint32_t* flags;    /* device, borrowed: per-cell flag words                      */
    int32_t   nx, ny, nz_local; /* cells of this partition                            */
    int32_t   z_halo;           /* ghost z planes per side, 0 (one device) or 1       */
    int64_t   pitch_y, pitch_z, pitch_q; /* in elements                               */
    int32_t   z_origin;         /* global z of local plane 0                          */
    int32_t   gnx, gny, gnz;    /* global box                                         */
    void*     wall_cache;       /* device, borrowed, may be NULL: x-face cache of the field in pop_out
                                   (nlbm_dense_wall_cache_build); lets the step kernels keep wall values
                                   of the cells at x = 0 and x = nx-1 without touching their rows   */
} nlbm_dense_desc;

int         nlbm_abi_version(void);
const char* nlbm_last_error(void);
/* number of CUDA devices visible, or -1 (error text in nlbm_last_error) */
int nlbm_device_count(void);

/* Fills pitch_y/pitch_z/pitch_q of `d` from nx, ny, nz_local, z_halo for an element
 * of elem_bytes (4|8) and returns the bytes one population FIELD of q components
 * needs in *pop_bytes and the flag array in *flag_bytes.  Host only.  This is where
 * the reference decides its pitch: dField_imp.h:67-87.                            */
int nlbm_dense_layout(nlbm_dense_desc* d, int q, int elem_bytes, size_t* pop_bytes, size_t* flag_bytes);

/* ---- x-face cache of a population field (optional, an optimisation of the step kernels) -----------------------------
 * Non-bulk cells are never updated (LbmTools.h:304), yet the thread that owns a wall cell at x = 0 or x = nx-1 together
 * with bulk cells stores 16 bytes per population and has to put the wall cell's present value of the OUTPUT field back.
 * Reading that value from its row costs one isolated DRAM access (a row activation) per row, population and side — as
 * many activations again as the whole streaming sweep (measured: 6-8 % of the iteration).  The cache holds those values
 * contiguously: cache[side][q][zm][y] = field[q][zm][y][side ? nx-1 : 0].  It stays valid as long as nobody but the step
 * kernels writes the field: build it once after the field was initialised / uploaded, rebuild it after any other write.
 * d->pop_out = the field, d->wall_cache = its cache (bytes from nlbm_dense_wall_cache_layout, 128-byte aligned).
 * With d->wall_cache == NULL in a step call the kernel reads the field itself: same results, slower.                  */
int nlbm_dense_wall_cache_layout(const nlbm_dense_desc* d, int q, int elem_bytes, size_t* bytes);
int nlbm_dense_wall_cache_build(const nlbm_dense_desc* d, int q, int elem_bytes, void* stream);

/* ---- problem set-up on the device (RunCavityTwoPop.cu:159-242) ------------------
 * geom: 0 lid-driven cavity; 1 cavity + solid sphere; 2 flow over sphere (x=0 inlet
 * treated as moving wall, other faces bounce-back).  sphere = {cx,cy,cz,R}
 * in global cells or NULL for the default (0.45nx, 0.55ny, 0.5nz, min(n)/5).
 * Writes the class bits of every plane including ghosts (mask bits cleared).       */
int nlbm_dense_classify(const nlbm_dense_desc* d, int geom, const double* sphere, void* stream);
/* Rebuilds the per-row summary and the cell map from the flag words (after the caller wrote them itself). */
/* Self-test of the two exact building blocks of that evaluation (synchronous, current device; used by tests/):
 * kind 0: float -> double by integer multiply-add against the conversion instruction for EVERY positive normal float;
 * kind 1: the shared-reciprocal division against IEEE division on n pseudo-random operand sets inside its guard.
 * *mismatches receives the number of differing results (must be 0).                                              */
int nlbm_selftest_exact(int kind, uint64_t n, uint64_t seed, uint64_t* mismatches);
int nlbm_dense_flags_commit(const nlbm_dense_desc* d, void* stream);
/* Flag words from a host-made classification (FieldBase::updateDeviceData of the CellType field, RunCavityTwoPop.cu:226-233):
 * `classes` (DEVICE pointer) holds one byte per cell, dense [nplanes][ny][nx], for the memory planes
 * [zm_first, zm_first + nplanes); every other plane and the row padding become `undefined`; wall bits are cleared;
 * summary and cell map are rebuilt.  4 x less to upload than flag words.                                      */
int nlbm_dense_flags_from_classes(const nlbm_dense_desc* d, const uint8_t* classes, int zm_first, int nplanes, void* stream);
/* LbmContainers::computeWallNghMask, LbmTools.h:344-376 (bit-exact).  q = 19|27.
 * Needs valid class bits in the ghost planes.  *d_bad (device int32, may be NULL) is
 * incremented for every bulk-cell neighbour outside the global domain.             */
int nlbm_dense_wall_mask(const nlbm_dense_desc* d, int q, int32_t* d_bad, void* stream);
/* Initial populations of BOTH semantics of the reference set-up: bulk t_k, bounceBack 0,
 * movingWall -6 t_k ulb (c_k . (1,0,0))   (RunCavityTwoPop.cu:168-206).  Writes d->pop_out. */
int nlbm_dense_init_pop_f32(const nlbm_dense_desc* d, int q, double ulb, void* stream);
int nlbm_dense_init_pop_f64(const nlbm_dense_desc* d, int q, double ulb, void* stream);

/* ---- THE HOT PATH: fused pull-stream + BGK collide, one iteration ---------------
 * LbmContainers::iteration (LbmTools.h:285-325) = pullStream (:99-168) + macroscopic
 * (:172-195) + collideBgkUnrolled (:199-282); D3Q27: apps/lbmMultiRes/{stream.h:5-49,
 * collide.h:286-354}.  Reads d->pop_in (+ghost planes), writes bulk cells of d->pop_out
 * in the z range selected by data_view.  omega as in Config.cpp:105-111.            */
int nlbm_d3q19_f32_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q19_f64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream);
/* store float / compute double — the reference sweep's "f/d" column */
int nlbm_d3q19_f32c64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q27_f32_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q27_f64_dense_step(const nlbm_dense_desc* d, double omega, int data_view, int opts, void* stream);

/* ---- fused iteration + halo update (one kernel does both, over peer memory) --------------------------------------
 * The reference runs, per iteration and device, [host sync -> 19 peer copies per direction -> kernel] (Occ::none) or the
 * INTERNAL kernel next to [sync -> copies -> BOUNDARY kernel] (Occ::standard; multiGpuGraph.cpp:120-143,304-352).  Here ONE
 * launch updates the whole partition, takes its two z-boundary planes first, stores the populations that cross each face
 * (D3Q19: 5, D3Q27: 9) straight into the neighbour's ghost plane of the field that corresponds to pop_out (a peer / CUDA-IPC
 * mapping; NVLink stores), and the last warp of a plane publishes `value` in the neighbour's flag word.  The caller
 * enqueues nlbm_flag_wait(my flag >= value of the previous iteration) before the next launch: the ghost planes an
 * iteration reads were written during the neighbours' previous iteration, so the exchange hides behind the interior.
 * The ghost planes of pop_in must be current at the first call (one ordinary halo update).  nz_local >= 2.
 * kind: 0 D3Q19 f32, 1 D3Q19 f64, 2 D3Q19 f32 store / f64 compute, 3 D3Q27 f32, 4 D3Q27 f64.                            */
typedef struct nlbm_peer_desc {
    void*     down_field;    /* lower neighbour's field corresponding to pop_out (device pointer valid HERE), or NULL */
    void*     up_field;      /* upper neighbour's, or NULL */
    int32_t   down_nz_local; /* slab heights of the neighbours (they fix where their ghost planes are) */
    int32_t   up_nz_local;
    uint32_t* down_flag;     /* word in the lower neighbour's memory: receives `value` when my plane 0 is in its upper ghost */
    uint32_t* up_flag;       /* word in the upper neighbour's memory: receives `value` when my top plane is in its lower ghost */
    uint32_t* counters;      /* 2 words of THIS device's memory, zero before the first call, owned by the caller */
    uint32_t  value;
} nlbm_peer_desc;
/* ---- several iterations without returning to the stream (small boxes) ----------------------------------------------------
 * `iterations` LBM iterations of a partition WITHOUT neighbours (d->z_halo == 0): iteration t reads d->pop_in when t is even,
 * else d->pop_out, and writes the other field — the two-field scheme of LbmIteration.h:56-61 without returning to the host (or to
 * the stream) in between.  The result is in pop_out when `iterations` is odd, else in pop_in.  d->wall_cache is pop_out's x-face
 * cache as in a step call, wall_cache_in the one of pop_in (both may be NULL).  kind as in nlbm_dense_step_push (0 d3q19_f32 ...
 * 4 d3q27_f64).  A box of a few hundred thousand cells iterates in ~10 us as a kernel of its own — launch, ramp-up and tail cost
 * as much as the work.  Default: a CHAIN of dependent launches (programmatic dependent launch, one per iteration): iteration t+1
 * is launched while t still runs, and a tile of plane z starts as soon as planes z-1, z, z+1 of the previous iteration are
 * complete (per-plane counters in device memory) — no grid ever waits for a whole grid, so launch gap, ramp-up and tail of
 * consecutive iterations overlap.  Only the first chip-load of planes of an iteration polls (NLBM_OPT_CHAIN_EARLY); the blocks
 * behind them are scheduled when the previous launch is over and pass griddepcontrol.wait at once.  A view of more than 4096
 * planes runs as `iterations` ordinary step launches.  (A single resident grid with a grid-wide barrier between iterations was
 * built and measured in round 2 — 16.5 us per 64^3 iteration against 11.8: profiles/r02k_multi64_*, r02l_* — and retired.)
 * Same results as `iterations` step calls, bit for bit.  Capturable into a CUDA graph.  The per-plane counters come from a pool the
 * library keeps per device (one slice per stream that issues chains, up to 16; one per captured chain, up to 48; allocated at
 * the first call outside a capture): a call that finds none issues `iterations` ordinary step launches instead.              */
int nlbm_dense_step_n(int kind, const nlbm_dense_desc* d, const void* wall_cache_in, double omega, int iterations, int opts, void* stream);

int nlbm_dense_step_push(int kind, const nlbm_dense_desc* d, const nlbm_peer_desc* peer, double omega, int opts, void* stream);

/* LbmContainers::computeRhoAndU, LbmTools.h:384-437 (D3Q19).  rho: [zm][y][x] with the
 * descriptor's pitches; u: 3 such planes sets, pitch_q apart.                        */
int nlbm_d3q19_f32_dense_rho_u(const nlbm_dense_desc* d, void* rho, void* u, void* stream);
int nlbm_d3q19_f64_dense_rho_u(const nlbm_dense_desc* d, void* rho, void* u, void* stream);

/* ---- halo update (dField::newHaloUpdate, dField_imp.h:341-421,548-641) ----------
 * The reference copies all q planes per direction with one cudaMemcpyPeerAsync each
 * (DataTransferContainer.h:38-55).  Here one launch moves only the populations that
 * cross the face (D3Q19: 5 of 19, D3Q27: 9 of 27; `lattice_q` = 0 moves all `ncomp`
 * components — "grid semantic", used for flags and generic fields).
 *   dir = +1: src's top boundary plane (local z = nz_local-1) -> dst's lower ghost plane
 *   dir = -1: src's bottom boundary plane (local z = 0)       -> dst's upper ghost plane
 * src/dst are device pointers to whole fields with the descriptors' layout; they may live
 * on different GPUs (peer access enabled by the caller, or a CUDA-IPC mapping): the
 * kernel then stores straight into the peer's HBM over NVLink.                       */
int nlbm_dense_halo_push(const nlbm_dense_desc* src_desc, const void* src_field, const nlbm_dense_desc* dst_desc,
                         void* dst_field, int elem_bytes, int ncomp, int lattice_q, int dir, void* stream);
/* Both faces of a partition in ONE launch, signalling included (replaces the same 2 x 19 cudaMemcpyPeerAsync + host syncs,
 * DataTransferContainer.h:38-55 / SynchronizationContainer.h:37-42): src's top plane goes into the lower ghost plane of the
 * neighbour above (up_field: its field through a peer / CUDA-IPC mapping, up_nz_local: its slab height), src's bottom plane
 * into the upper ghost plane of the neighbour below; when all stores of the launch are out, `value` is published in both
 * neighbours' flag words (nlbm_flag_wait2 on their side).  Either neighbour may be NULL.  counter: one word of device
 * memory on THIS GPU, zero before the first call, owned by the caller.                                                  */
int nlbm_dense_halo_push2(const nlbm_dense_desc* src_desc, const void* src_field, void* up_field, int32_t up_nz_local, uint32_t* up_flag,
                          void* down_field, int32_t down_nz_local, uint32_t* down_flag, uint32_t* counter, uint32_t value,
                          int elem_bytes, int ncomp, int lattice_q, void* stream);
/* Staged variant for transports that cannot map peer memory (NCCL send/recv):
 * pack the crossing populations of one boundary plane into a contiguous buffer and
 * unpack such a buffer into a ghost plane.  Returns the byte count in *bytes.        */
int nlbm_dense_halo_pack(const nlbm_dense_desc* d, const void* field, int elem_bytes, int ncomp, int lattice_q, int dir,
                         void* buffer, size_t* bytes, void* stream);
int nlbm_dense_halo_unpack(const nlbm_dense_desc* d, void* field, int elem_bytes, int ncomp, int lattice_q, int dir,
                           const void* buffer, void* stream);


/* ================================================================================================================
 * Block-sparse (bGrid) partitions: 8 x 8 x 8-cell blocks (Neon::bGrid = StaticBlock<8,8,8>, libNeonDomain/include/Neon/
 * domain/bGrid.h:5).  Replaces what the reference kernel receives by value: bSpan {firstDataBlockOffset, dataView}
 * (bSpan.h:47-49) and bPartition members (bPartition.h:154-160: mem, cardinality, blockConnectivity, mask, origin).
 *
 * Memory layout (SoA per population; every block's tile of a population is one contiguous 2 KB-aligned run):
 *   pop[q][blk][z][y][x]   element offset (q * n_blocks_alloc + blk) * 512 + z*64 + y*8 + x
 *   flags[blk][z][y][x]    the dense flag word; cells that are not active carry class NLBM_UNDEFINED (this replaces the
 *                          per-block active bit mask, StaticBlock.h:47-103)
 *   info[blk][32]          words 0..26: neighbour block ids, index (dx+1) + 3*(dy+1) + 9*(dz+1) (bPartition_imp.h:194-198),
 *                          NLBM_NO_BLOCK where there is none; words 27..29: global origin x, y, z of the block; 30..31 spare
 * Block order inside a partition: [0, n_blocks) local blocks, sorted so that the blocks of the lowest block layer come
 * first (n_down of them) and those of the highest layer last (n_up); [n_blocks, n_blocks_alloc) ghost blocks (copies of
 * the facing layers of the z-neighbours: first the n_ghost_down blocks below, then those above).  (Reference:
 * [blk][q][z][y][x] with 32-bit offsets — overflows beyond 2^32/19 cells per device, bPartition_imp.h:97-107.)          */
#define NLBM_NO_BLOCK 0xFFFFFFFFu
typedef struct nlbm_block_desc {
    void*           pop_in;   /* device, borrowed */
    void*           pop_out;  /* device, borrowed */
    uint32_t*       flags;    /* device, borrowed */
    const uint32_t* info;     /* device, borrowed */
    uint32_t        n_blocks;        /* local blocks (the ones this partition updates) */
    uint32_t        n_blocks_alloc;  /* local + ghost blocks */
    uint32_t        n_down, n_up;    /* local blocks in the lowest / highest block layer (BOUNDARY view), 0 if unsplit */
    int32_t         gnx, gny, gnz;   /* global box in cells */
} nlbm_block_desc;

/* Set-up on the device, as for dense partitions.  active_mask: 16 words per block (bit z*64+y*8+x, StaticBlock.h:47-103)
 * or NULL = every in-domain cell of every block is active.  classify covers local and ghost blocks.                       */
int nlbm_block_classify(const nlbm_block_desc* d, int geom, const double* sphere, const uint32_t* active_mask, void* stream);
int nlbm_block_wall_mask(const nlbm_block_desc* d, int q, int32_t* d_bad, void* stream);
int nlbm_block_init_pop_f32(const nlbm_block_desc* d, int q, double ulb, void* stream);
int nlbm_block_init_pop_f64(const nlbm_block_desc* d, int q, double ulb, void* stream);

/* THE HOT PATH on bGrid: LbmContainers::iteration (LbmTools.h:285-325) over the blocks of data_view
 * (STANDARD: all local blocks; BOUNDARY: the n_down + n_up blocks of the outer layers; INTERNAL: the rest).              */
int nlbm_d3q19_f32_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q19_f64_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q27_f32_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream);
int nlbm_d3q27_f64_block_step(const nlbm_block_desc* d, double omega, int data_view, int opts, void* stream);

/* Halo update of block-sparse partitions (bField::newHaloUpdate, bField_imp.h:173-332 — which ignores the cardinality and
 * copies whole blocks; measured NaN upstream for Q = 19).  Here: for the populations that cross the face
 * (lattice_q = 19 | 27; 0 = all ncomp components) only the facing z-slice (64 cells) of every boundary block moves.
 *   dir = +1: z-slice 7 of src's n_up highest-layer blocks -> the same slice of dst's ghost-down blocks [dst_first_ghost ..)
 *   dir = -1: z-slice 0 of src's n_down lowest-layer blocks -> the same slice of dst's ghost-up blocks
 * The i-th boundary block of src corresponds to the i-th ghost block of dst (both sorted by block y, x).                   */
int nlbm_block_halo_push(const nlbm_block_desc* src_desc, const void* src_field, const nlbm_block_desc* dst_desc, void* dst_field,
                         uint32_t dst_first_ghost, int elem_bytes, int ncomp, int lattice_q, int dir, void* stream);

/* Device-side ordering for the peer-store halo transport between PROCESSES (one per GPU): the reference orders its
 * cudaMemcpyPeerAsync copies with events of one process plus host-blocking syncs (SynchronizationContainer.h:37-42);
 * across processes a counter word in the receiver's memory replaces the event.
 *   nlbm_flag_signal: enqueue "*flag = value" after everything already in `stream` (system-wide visibility);
 *                     `flag` may be a CUDA-IPC / peer mapping of memory on the neighbouring GPU.
 *   nlbm_flag_wait  : enqueue a wait until *flag >= value (wrap-safe).  If timeout_ms elapse first, *d_err (device int32,
 *                     may be NULL) is incremented and the kernel TRAPS: the work queued behind the wait would read a ghost
 *                     plane that never arrived, so the failure is made fatal for the process (every later CUDA call
 *                     returns an error) instead of letting the stream continue on stale data.
 *   nlbm_flag_wait2 : the same for two flag words in one launch (a partition has at most two z-neighbours); either may
 *                     be NULL.                                                                                      */
/* CUDA-IPC plumbing for the peer-store transport between processes.  export: handle of the ALLOCATION that contains `ptr`
 * and ptr's byte offset inside it; import (in another process, with ITS device current): maps that allocation and returns
 * its base — peer access to the exporting device is enabled on demand.  A handle must be imported once per process.       */
int nlbm_ipc_export(const void* ptr, unsigned char* handle64, uint64_t* offset);
int nlbm_ipc_import(const unsigned char* handle64, void** base);
int nlbm_ipc_close(void* base);
/* Enables direct access from the current device to `peer_device` (the reference does this for every device pair when
 * the Backend is built, libNeonSet/src/set/DevSet.cpp:80-100).  Idempotent.                                          */
int nlbm_enable_peer_access(int peer_device);
int nlbm_flag_signal(uint32_t* flag, uint32_t value, void* stream);
int nlbm_flag_wait(const uint32_t* flag, uint32_t value, uint32_t timeout_ms, int32_t* d_err, void* stream);
int nlbm_flag_wait2(const uint32_t* flag_a, const uint32_t* flag_b, uint32_t value, uint32_t timeout_ms, int32_t* d_err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEON_LBM_H */
