// Host-side derived task lists for the device executor ("inspector-executor" split).
//
// Input  : the reference inspector's arrays exactly as cholesky_left_par_05 receives them
//          (cholesky/parallel_PB_Cholesky_05.h:27-39; layout SURVEY.md Appendix A).
// Output : per dependency step, the batches the CUDA kernels consume:
//            * small supernodes (width <= SMALL_W) factored by one warp each,
//            * block columns (<= NB wide) of the remaining supernodes: POTRF + TRSM tasks,
//            * update pairs (target supernode, descendant) with (lb, ndrow1, nSupRs) — what the reference
//              finds by scanning the descendant's rows (:137-152) for every supernode returned by ereach_sn
//              (common/Reach.h:112) — split into DMMA tile tasks and warp-FMA tasks.
// Nothing here touches numeric values.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <string>
#include <new>
#include <utility>

namespace parsy {

// Allocator of the planner's long lists.  First-touch page faults are the largest single cost of building a plan for
// a 2-D problem (about 100 MB of lists; ~1.7 us per 4 KB page on the hosts measured, three times the cost of the writes
// themselves), so lists of 4 MB and more are mapped directly, 2 MB-aligned, with MADV_HUGEPAGE; smaller ones come from
// malloc.  resize(n) leaves new elements uninitialised (no second pass over the list): every user fills what it sizes.
void* big_alloc(size_t bytes);
void big_free(void* p, size_t bytes);
template <class T> struct BigAlloc {
  using value_type = T;
  BigAlloc() = default;
  template <class U> BigAlloc(const BigAlloc<U>&) {}
  T* allocate(size_t n) { return (T*)big_alloc(n * sizeof(T)); }
  void deallocate(T* p, size_t n) { big_free(p, n * sizeof(T)); }
  template <class U> void construct(U* p) { ::new ((void*)p) U; }
  template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
  template <class U> bool operator==(const BigAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const BigAlloc<U>&) const { return false; }
};
template <class T> using BigVec = std::vector<T, BigAlloc<T>>;

constexpr int SMALL_W = 32;       // widest supernode handled by the warp-cooperative path
constexpr int SMALL_R = 1024;     // ... and its largest row count
constexpr int64_t SMALL_WORK = 100000;   // ... and the bound on (rows below) * width^2 one warp is asked to do
constexpr int SMALL_W_NARROW = 8;        // supernodes / pairs at most this wide run in the low-register kernels
constexpr int FAR_STEPS = 8;      // see Step::upd
constexpr int NB_MAX = 128;       // largest block-column width (POTRF tile held in shared memory)

enum GemmFlags : int32_t {
  GF_LOWER = 1,      // only elements with row >= col (pair-local) are applied (SYRK part of the trapezoid)
  GF_ATOMIC = 2,     // target shared with other tasks of the launch: red.global.add.f64
  GF_OVERWRITE = 4,  // C = A*B' (TRSM through the inverse of the diagonal block), else C -= A*B'
  GF_B_LINV = 8,     // B operand lives in the inverse-diagonal-block store, not in the factor
};

// One GEMM-shaped task: C[rel(i), rel(j)] (-)= sum_k A[i,k] * B[j,k], i < M, j < N, k < K.
struct GemmTask {
  int64_t a_off;     // offset of A(0,0) in lValues
  int64_t b_off;     // offset of B(0,0) in lValues (or in the Linv store with GF_B_LINV)
  int64_t c_off;     // offset of the target panel's (0,0)
  int64_t rel_off;   // offset into the relative-index array, -1 = identity map
  int32_t lda, ldb, ldc;
  int32_t M, N, K;
  int32_t flags;
  int32_t tile0;     // first tile of this task within its launch
};
static_assert(sizeof(GemmTask) == 64, "GemmTask layout");

// Row chunk of a small pair, one warp each.
struct SmallTask {
  int32_t pair;      // index into the GemmTask array
  int32_t row0;      // first pair-local row of the chunk
  int32_t nrows;
  int32_t pad;
};

struct SupInfo {
  int64_t rowptr;    // Li_ptr[col0]: start of the row list in lR
  int64_t valptr;    // lC[col0]: start of the panel in lValues
  int32_t col0, w, r;
  int32_t flags;     // 1 = small path
};
static_assert(sizeof(SupInfo) == 32, "SupInfo layout");

struct BlockTask {
  int32_t sup;       // supernode
  int32_t j0;        // first local column of the block column
  int32_t nb;        // its width (<= NB)
  int32_t slot;      // index of its inverse-diagonal-block slot
  int32_t tile0;     // solve kernels: first row tile of this task within the launch
  int32_t pad[3];
};
static_assert(sizeof(BlockTask) == 32, "BlockTask layout");

// Pair before classification (also used by the tests to compare against ereach_sn).
// ---- single-launch dataflow sweeps -------------------------------------------------------------------
// A "node" is what gets solved as one unit: a narrow supernode or one block column.  A "solve task" is what one
// warp (narrow supernode) or one CTA (128-row slice of a block column's rows below the diagonal block) does.
constexpr int SOLVE_TILE_ROWS = 64;     // rows a sweep CTA holds in registers at a time
constexpr int SOLVE_TASK_ROWS = 256;    // rows of one solve task (4 register tiles share one diagonal-block solve)
constexpr int SOLVE_TALL_ROWS = 64;     // a narrow supernode with more rows below its diagonal block than this, and at
constexpr int SOLVE_TALL_WORK = 768;    // least this many panel entries below it, gets a CTA of its own in the sweeps
constexpr int SOLVE_CRITICAL_TASKS = 2; // leading one-tile tasks of a block column: the rows of the NEXT block column
struct SolveTask {
  int32_t sup;        // supernode
  int32_t node;       // node this task belongs to (the one whose solution it consumes)
  int32_t j0, nb;     // block column (narrow supernode: 0, width)
  int32_t slot;       // inverse-diagonal-block slot (-1 for narrow supernodes)
  int32_t row0;       // first local row of the slice (narrow: width)
  int32_t nrows;      // rows in the slice (may be 0: a block column with nothing below it)
  int32_t first;      // 1 = this task publishes the node's solution (slice 0)
  int32_t tgt_begin, tgt_end;   // range in solve_targets: nodes whose unknowns this task's rows touch
  int32_t tile_tgt[4];          // block-column tasks: end of the targets of each SOLVE_TILE_ROWS-row tile (tile k owns
                                // [k ? tile_tgt[k-1] : tgt_begin, tile_tgt[k])): the forward sweep publishes tile by tile
  int32_t pad[2];
  // copies of the supernode's descriptor and of the node's forward-sweep input count: a narrow task is a chain of
  // dependent memory round trips, and these save two of them (sup[T.sup], need[T.node])
  int64_t rowptr, valptr;
  int32_t col0, r, need, pad2;
};
static_assert(sizeof(SolveTask) == 96, "SolveTask layout");
static_assert(SOLVE_TASK_ROWS == 4 * SOLVE_TILE_ROWS, "SolveTask::tile_tgt holds four tiles");
struct SolveCta {     // work of one CTA of the sweep kernel
  int32_t kind;       // 0 = up to 8 narrow supernodes (one per warp), 1 = one slice of a block column,
                      // 2 = one narrow supernode with a long panel, all eight warps on its rows
  int32_t first;      // index of the first SolveTask
  int32_t count;
  int32_t pad;
};

struct PairDesc {
  int32_t tgt, src;  // supernodes
  int32_t lb;        // first row (index in src's row list) that falls inside tgt's columns
  int32_t nd1;       // number of src rows inside tgt's columns           (ndrow1)
  int32_t m;         // rows from lb to the end of src                     (nSupRs = ndrow1 + ndrow3)
};

struct Range { int32_t begin = 0, end = 0; int32_t size() const { return end - begin; } };

// Update work of one step, split by when its target is needed.
struct UpdGroup {
  Range u128;          // gemm_tasks, 128x128 DMMA tiles
  int32_t tiles128 = 0;
  Range u64;           // gemm_tasks, 64x64 DMMA tiles
  int32_t tiles64 = 0;
  Range u32;           // gemm_tasks, 32x32 DMMA tiles (launches that cannot fill the GPU with larger tiles)
  int32_t tiles32 = 0;
  Range small_pairs;   // gemm_tasks of the warp-FMA pairs (addressed through small_tasks)
  Range small;         // small_tasks (row chunks); the first small_narrow have K <= 4
  int32_t small_narrow = 0;
};

struct Step {
  int32_t hlevel = 0;
  Range small_sup;   // into small_list, sorted by width; the first small_narrow have width <= SMALL_W_NARROW
  int32_t small_narrow = 0;
  Range blocks;      // into block_tasks
  int32_t blocks_owned = 0;  // leading block tasks this rank factors itself (distributed top: the owner's; else all)
  Range trsm;        // into gemm_tasks (T128, GF_OVERWRITE|GF_B_LINV)
  int32_t trsm_tiles = 0;
  int32_t trsm_tm = 64;      // row-tile height of this step's TRSM launch (64, 32 or 16)
  UpdGroup upd[3];   // [0] targets factored in the next step ("A"), [1] the others ("R"); distributed top only:
                     // [2] targets at least FAR_STEPS steps away ("far", own stream)
  int8_t upd_remote[3] = {0, 0, 0};   // distributed top: the group reads a panel another rank factors (wait for the broadcast)
  int32_t solve_tiles = 0;   // row tiles of the block tasks (forward / backward sweeps)
  int32_t max_nb = 0;        // widest block column in this step
};

struct PlanOptions {
  int nb = 128;
  bool ignore_hlevels = false;
  // multi-GPU sharding: the bottom of the tree (every H-level but the last `top_levels`) is a forest
  // of subtrees, dealt to the ranks in contiguous column order balanced by their flops; the top separators are shared.
  //   phase 0: everything (single GPU)
  //   phase 1: the subtrees `rank` owns AND every update they push into the top separators (fan-in: each rank
  //            accumulates the contributions of its own descendants in its copy of the top panels; a sum over the
  //            ranks then delivers A_top - sum of all bottom updates)
  //   phase 2: the top separators and the updates between them
  int phase = 0, rank = 0, world = 1, top_levels = 1;
  // phase 2: 1 = the top separators are distributed (1-D block-cyclic by block column: the owner of a block column
  // applies every update into it, factors it (POTRF + TRSM) and broadcasts the finished panel), 0 = the top is
  // computed redundantly by every rank
  int top_distributed = 1;
  // consecutive block columns of the top given to one rank before moving on to the next (block-cyclic with this block
  // size): inside such a run the chain POTRF -> TRSM -> next-column update stays on one GPU and the panel broadcast
  // leaves the critical path
  int top_chunk = 4;
};

struct Plan {
  int32_t n = 0, nsuper = 0, nlevels = 0;
  int64_t xsize = 0, ssize = 0, nnzA = 0;
  int nb = 128;
  BigVec<SupInfo> sup;
  std::vector<int32_t> small_list;
  std::vector<BlockTask> block_tasks;
  BigVec<GemmTask> gemm_tasks;
  BigVec<SmallTask> small_tasks;
  std::vector<Step> steps;
  std::vector<int32_t> hlevel_first_step;   // nlevels+1
  // relative indices: rel_src_row[e] = global row index to look up, rel_tgt[e] = target supernode;
  // filled on the device from (pair prefix) — here only the per-pair prefix and descriptors
  BigVec<PairDesc> pairs;                   // all pairs, in GemmTask order for the first pairs.size() real pairs
  BigVec<int64_t> rel_prefix;               // per real pair with a map: start in the rel array (size npairs+1)
  BigVec<int32_t> rel_pair_src;             // per rel-pair: source supernode
  BigVec<int32_t> rel_pair_tgt;             // per rel-pair: target supernode
  BigVec<int32_t> rel_pair_lb;              // per rel-pair: lb
  int64_t rel_entries = 0;
  std::vector<int32_t> owner;               // per supernode: owning rank, -1 = shared top (world > 1 only)
  std::vector<int32_t> node_owner;          // per node (narrow supernode / block column) of the top: owning rank
  std::vector<int32_t> bcast_ptr;           // per step: range in bcast (phase 2, distributed top)
  std::vector<int64_t> bcast;               // triples (owner, begin, end) in doubles: panels the owner broadcasts once the
                                            // step's block columns are factored
  std::vector<int32_t> bcast_shape;         // per broadcast: (j0, rows of the supernode, block-column width) — rows above
                                            // j0 of a block column are structural zeros and need not travel
  std::vector<BlockTask> invert_tasks;      // phase 2, distributed top: block columns factored by OTHER ranks; their inverse
                                            // diagonal blocks (diagonal solves of the sweeps) are rebuilt locally
  std::vector<int64_t> zero_runs;           // pairs (begin, end) in doubles: what a factorization zeroes first
  std::vector<int64_t> top_runs;            // pairs (begin, end): panels of the shared top (summed over the ranks)
  std::vector<int32_t> col_runs;            // pairs (begin, end): columns solved by this plan's sweeps
  std::vector<uint8_t> skip_assemble;       // per supernode: 1 = entries of A are NOT scattered by this plan
  int32_t first_top_step = 0;
  // dataflow sweeps
  int32_t n_nodes = 0;
  BigVec<SolveTask> solve_tasks;           // forward order (dependency steps ascending)
  std::vector<SolveCta> solve_ctas;        // forward order; the backward sweep walks it in reverse
  int32_t n_narrow_prefix_ctas = 0;        // leading CTAs (all of kind 0) that come before the first block-column task:
                                           // the leaf region of the tree, run by the light narrow-only sweep kernels
  BigVec<int32_t> solve_targets;           // node ids
  std::vector<int32_t> node_need;          // forward: number of tasks that add into the node's right-hand side
  std::vector<int32_t> node_tiles;         // backward: number of slices of the node (narrow supernodes: 1)
  int32_t n_slots = 0;
  int64_t n_pairs = 0, n_pairs_small = 0, n_pairs_tiled = 0, n_block_cols = 0;
  double flops_potrf = 0, flops_trsm = 0, flops_update = 0, bytes_solve = 0;
  // algorithmic flops per kernel class: 0 factor_small, 1 potrf_block, 2 trsm tiles, 3 update tiles 128, 4 update
  // tiles 64, 5 update_small
  double class_flops[6] = {0, 0, 0, 0, 0, 0};
  std::string error;
};

// Returns 0 on success, PARSY_CUDA_ERR_* otherwise (message in plan.error).
// Number of places where the sweep task lists violate the ordering the spinning kernels rely on (0 = deadlock-free).
int64_t sweep_order_violations(const Plan& P);
// 64-bit FNV-1a over every list of the plan the executor reads (task lists, step table, sweep plan, ownership, runs):
// two plans with the same digest launch the same work on the same operands.
uint64_t plan_digest(const Plan& P);
int build_plan(Plan& plan, int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet,
               int supNo, const int* aTree, const int* col2Sup, int nLevels, const int* levelPtr, const int* parPtr,
               const int* partition, const PlanOptions& opt);

// Descendant pairs only (no schedule needed): used by tests against ereach_sn and by the prune-set entry point.
void enumerate_pairs(std::vector<PairDesc>& out, int supNo, const int* blockSet, const size_t* Li_ptr, const int* lR,
                     const int* col2Sup);

}  // namespace parsy
