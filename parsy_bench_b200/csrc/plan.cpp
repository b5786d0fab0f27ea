// Planner: turns the inspector's arrays + LBC schedule into per-step kernel batches. See plan.h.
#include "plan.h"
#include "../../include/parsy_cuda.h"
#include <algorithm>
#include <climits>
#include <cstring>
#include <chrono>
#include <thread>
#include <exception>
#include <stdexcept>
#include <cstdio>
#include <cstdlib>
#include <sys/mman.h>

namespace parsy {

void* big_alloc(size_t bytes) {
  constexpr size_t HP = (size_t)2 << 20;
  if (bytes < 2 * HP) { void* p = malloc(std::max<size_t>(bytes, 1)); if (!p) throw std::bad_alloc(); return p; }
  const size_t len = (bytes + HP - 1) / HP * HP;
  char* raw = (char*)mmap(nullptr, len + HP, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (raw == (char*)MAP_FAILED) throw std::bad_alloc();
  char* p = (char*)(((uintptr_t)raw + HP - 1) / HP * HP);
  if (p > raw) munmap(raw, (size_t)(p - raw));                       // give the unaligned head and the tail back
  if (p + len < raw + len + HP) munmap(p + len, (size_t)(raw + len + HP - (p + len)));
  madvise(p, len, MADV_HUGEPAGE);                                    // advisory: without THP the list is simply 4 KB-paged
  return p;
}
void big_free(void* p, size_t bytes) {
  constexpr size_t HP = (size_t)2 << 20;
  if (bytes < 2 * HP) { free(p); return; }
  munmap(p, (bytes + HP - 1) / HP * HP);
}

// For descendant d, every maximal run of its off-diagonal rows that falls into the columns of one supernode t is one
// update pair (t, d).  The reference finds the same [lb, ub] by a linear scan for each d returned by ereach_sn
// (parallel_PB_Cholesky_05.h:137-152); rows are sorted, so runs are contiguous.  Appends the pairs of d0 <= d < d1.
template <class Vec> static void pairs_of_range(Vec& out, int d0, int d1, const int* blockSet, const size_t* Li_ptr,
                           const int* lR, const int* col2Sup, int64_t* count_of_src) {
  for (int d = d0; d < d1; ++d) {
    const int col0 = blockSet[d], w = blockSet[d + 1] - col0;
    const size_t rp = Li_ptr[col0];
    const int r = (int)(Li_ptr[blockSet[d + 1]] - rp);
    const size_t before = out.size();
    int i = w;
    while (i < r) {
      const int t = col2Sup[lR[rp + i]];
      int e = i + 1;
      while (e < r && col2Sup[lR[rp + e]] == t) ++e;
      out.push_back(PairDesc{t, d, i, e - i, r - i});
      i = e;
    }
    if (count_of_src) count_of_src[d] = (int64_t)(out.size() - before);
  }
}

void enumerate_pairs(std::vector<PairDesc>& out, int supNo, const int* blockSet, const size_t* Li_ptr, const int* lR,
                     const int* col2Sup) {
  out.clear();
  pairs_of_range(out, 0, supNo, blockSet, Li_ptr, lR, col2Sup, nullptr);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Planner threads: PARSY_PLAN_THREADS, default min(8, hardware threads).  The lists are built by index ranges that do not
// overlap, in an order fixed by prefix sums, so the plan does not depend on the number of threads.
static int plan_threads() {
  static const int nt = [] {
    const char* e = getenv("PARSY_PLAN_THREADS");
    int v = e ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, std::min(v, 64));
  }();
  return nt;
}
// fn(begin, end, part) over `parts` contiguous pieces of [0, n); the caller's thread takes the last piece
// An exception in any piece (std::bad_alloc, in practice) is rethrown in the caller once every thread has finished.
template <class F> static void par_ranges(size_t n, int parts, F fn) {
  parts = (int)std::max<size_t>(1, std::min<size_t>(parts, n / 4096 + 1));
  if (parts == 1) { fn((size_t)0, n, 0); return; }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(parts);
  auto piece = [&](int k) {
    try { fn(n * k / parts, n * (k + 1) / parts, k); } catch (...) { err[k] = std::current_exception(); }
  };
  try {
    th.reserve(parts - 1);
    for (int k = 0; k + 1 < parts; ++k) th.emplace_back(piece, k);
  } catch (...) {                                   // no more threads to be had: the caller does the rest itself
    for (int k = (int)th.size(); k + 1 < parts; ++k) piece(k);
  }
  piece(parts - 1);
  for (auto& t : th) t.join();
  for (int k = 0; k < parts; ++k) if (err[k]) std::rethrow_exception(err[k]);
}
// tiles (mi, ni) of a TM x TN grid over the lower trapezoid: column tile ni needs row tiles from (ni*TN)/TM on
static inline int lower_tiles(int M, int N, int TM, int TN) {
  const int MT = cdiv(M, TM), NT = cdiv(N, TN);
  int cnt = 0;
  for (int ni = 0; ni < NT; ++ni) cnt += MT - (ni * TN) / TM;
  return cnt;
}

static int build_plan_impl(Plan& P, int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet,
                           int supNo, const int* aTree, const int* col2Sup, int nLevels, const int* levelPtr,
                           const int* parPtr, const int* partition, const PlanOptions& opt);

// The planner allocates a few hundred bytes per supernode and pair on the host; running out of memory is reported, not
// thrown through the C ABI.
int build_plan(Plan& P, int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet, int supNo,
               const int* aTree, const int* col2Sup, int nLevels, const int* levelPtr, const int* parPtr,
               const int* partition, const PlanOptions& opt) {
  try {
    return build_plan_impl(P, n, lC, lR, Li_ptr, blockSet, supNo, aTree, col2Sup, nLevels, levelPtr, parPtr, partition, opt);
  } catch (const std::bad_alloc&) {
    P.error = "out of host memory while building the plan";
    return PARSY_CUDA_ERR_NO_MEMORY;
  } catch (const std::exception& e) {
    P.error = std::string("planner failed: ") + e.what();
    return PARSY_CUDA_ERR_NO_MEMORY;
  }
}

static int build_plan_impl(Plan& P, int n, const size_t* lC, const int* lR, const size_t* Li_ptr, const int* blockSet,
                           int supNo, const int* aTree, const int* col2Sup, int nLevels, const int* levelPtr,
                           const int* parPtr, const int* partition, const PlanOptions& opt) {
  const bool timing_on = getenv("PARSY_PLAN_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlast = tnow();
  auto lap = [&](const char* what) { if (timing_on) { const double t = tnow(); fprintf(stderr, "[plan] %-28s %.1f ms\n", what, (t - tlast) * 1e3); tlast = t; } };
  if (n < 0 || supNo < 0 || !lC || !lR || !Li_ptr || !blockSet || !col2Sup || !levelPtr || !parPtr ||
      !partition || nLevels < 0) {
    P.error = "NULL or negative argument";
    return PARSY_CUDA_ERR_BAD_ARG;
  }
  const int NB = opt.nb <= 0 ? 128 : opt.nb;
  if (NB > NB_MAX || NB < 8 || (NB % 8) != 0) { P.error = "block_cols must be a multiple of 8 in [8,128]"; return PARSY_CUDA_ERR_BAD_ARG; }
  P.n = n; P.nsuper = supNo; P.nlevels = nLevels; P.nb = NB;
  P.sup.resize(supNo);
  if (supNo > 0 && (blockSet[0] != 0 || blockSet[supNo] != n)) { P.error = "blockSet does not cover 0..n"; return PARSY_CUDA_ERR_BAD_ARG; }
  int64_t xs = 0, ss = 0;
  for (int s = 0; s < supNo; ++s) {
    SupInfo& I = P.sup[s];
    I.col0 = blockSet[s]; I.w = blockSet[s + 1] - blockSet[s];
    if (I.w <= 0) { P.error = "empty supernode"; return PARSY_CUDA_ERR_BAD_ARG; }
    I.rowptr = (int64_t)Li_ptr[I.col0];
    I.r = (int)(Li_ptr[blockSet[s + 1]] - Li_ptr[I.col0]);
    I.valptr = (int64_t)lC[I.col0];
    if (I.r < I.w) { P.error = "supernode with fewer rows than columns"; return PARSY_CUDA_ERR_BAD_ARG; }
    I.flags = (I.w <= SMALL_W && I.r <= SMALL_R && (int64_t)(I.r - I.w) * I.w * I.w <= SMALL_WORK) ? 1 : 0;
    xs = std::max<int64_t>(xs, I.valptr + (int64_t)I.w * I.r);
    ss = std::max<int64_t>(ss, I.rowptr + I.r);
    const double w = I.w, r = I.r;
    P.flops_potrf += w * w * w / 3.0;
    P.flops_trsm += w * w * (r - w);
    P.bytes_solve += 8.0 * (w * (w + 1) / 2 + w * (r - w));
  }
  P.xsize = xs; P.ssize = ss;
  P.bytes_solve += 4.0 * (double)ss + 16.0 * (double)n;

  lap("supernode table");
  // ---- schedule: validate, then assign dependency steps --------------------------------------------
  std::vector<int32_t> hl(supNo, -1), part(supNo, -1), pos(supNo, -1);
  std::vector<int32_t> order; order.reserve(supNo);
  for (int H = 0; H < nLevels; ++H)
    for (int j1 = levelPtr[H]; j1 < levelPtr[H + 1]; ++j1)
      for (int k1 = parPtr[j1]; k1 < parPtr[j1 + 1]; ++k1) {
        const int s = partition[k1];
        if (s < 0 || s >= supNo || hl[s] >= 0) { P.error = "schedule lists a supernode twice or out of range"; return PARSY_CUDA_ERR_BAD_SCHEDULE; }
        hl[s] = H; part[s] = j1; pos[s] = (int)order.size(); order.push_back(s);
      }
  if ((int)order.size() != supNo) { P.error = "schedule does not cover every supernode"; return PARSY_CUDA_ERR_BAD_SCHEDULE; }
  // sharded plans with a distributed top: every top supernode takes the block-column path (its block columns are the
  // units of ownership); set for both phases, so that their node numbering agrees
  const int first_top_level = std::max(0, nLevels - std::max(1, opt.top_levels));
  if (opt.world > 1 && opt.top_distributed)
    for (int s = 0; s < supNo; ++s) if (hl[s] >= first_top_level) P.sup[s].flags = 0;
  // every update pair (target, descendant) must run the descendant first: an earlier H-level, or earlier in
  // the same w-partition (SURVEY.md Appendix E legality condition)
  lap("schedule arrays");
  // the row scan, by ranges of descendants on several threads; each piece also checks its pairs (the first offence in
  // pair order is the one reported, as a serial pass would)
  std::vector<int64_t> src_ptr(supNo + 1, 0);
  {
    const int parts = plan_threads();
    std::vector<BigVec<PairDesc>> loc(parts);
    std::vector<int> bad(parts, 0);
    par_ranges((size_t)supNo, parts, [&](size_t d0, size_t d1, int k) {
      loc[k].reserve((d1 - d0) * 3);
      // every stored row index is looked up in col2Sup (here) and in the panels (device): refuse what lies outside
      if (d1 > d0)
        for (size_t e2 = Li_ptr[blockSet[d0]]; e2 < Li_ptr[blockSet[d1]]; ++e2)
          if ((unsigned)lR[e2] >= (unsigned)n) { bad[k] = 3; return; }
      pairs_of_range(loc[k], (int)d0, (int)d1, blockSet, Li_ptr, lR, col2Sup, src_ptr.data() + 1);
      for (const PairDesc& q : loc[k]) {
        if (q.tgt <= q.src || q.tgt >= supNo) { bad[k] = 1; break; }
        const bool ok = hl[q.src] < hl[q.tgt] || (hl[q.src] == hl[q.tgt] && part[q.src] == part[q.tgt] && pos[q.src] < pos[q.tgt]);
        if (!ok) { bad[k] = 2; break; }
      }
    });
    for (int k = 0; k < parts; ++k) {
      if (bad[k] == 3) { P.error = "row index outside the matrix"; return PARSY_CUDA_ERR_BAD_ARG; }
      if (bad[k] == 1) { P.error = "row structure is not lower triangular"; return PARSY_CUDA_ERR_BAD_ARG; }
      if (bad[k] == 2) { P.error = "schedule runs a supernode before one of its descendants"; return PARSY_CUDA_ERR_BAD_SCHEDULE; }
    }
    std::vector<size_t> base(parts + 1, 0);
    for (int k = 0; k < parts; ++k) base[k + 1] = base[k] + loc[k].size();
    P.pairs.resize(base[parts]);
    par_ranges((size_t)parts, parts, [&](size_t k0, size_t k1, int) {
      for (size_t k = k0; k < k1; ++k) if (!loc[k].empty()) memcpy(P.pairs.data() + base[k], loc[k].data(), loc[k].size() * sizeof(PairDesc));
    });
  }
  P.n_pairs = (int64_t)P.pairs.size();
  for (int s = 0; s < supNo; ++s) src_ptr[s + 1] += src_ptr[s];
  lap("pair scan");
  std::vector<int32_t> nblk(supNo), step0(supNo, 0), need(supNo, 0);
  for (int s = 0; s < supNo; ++s) nblk[s] = P.sup[s].flags ? 1 : cdiv(P.sup[s].w, NB);
  P.hlevel_first_step.assign(nLevels + 1, 0);
  int nsteps = 0;
  {
    int idx = 0;
    for (int H = 0; H < nLevels; ++H) {
      const int base = opt.ignore_hlevels ? 0 : nsteps;
      P.hlevel_first_step[H] = nsteps;
      int cnt = 0;
      for (int j1 = levelPtr[H]; j1 < levelPtr[H + 1]; ++j1) cnt += parPtr[j1 + 1] - parPtr[j1];
      for (int q = 0; q < cnt; ++q, ++idx) {
        const int s = order[idx];
        step0[s] = std::max(base, need[s]);
        const int last = step0[s] + nblk[s] - 1;
        nsteps = std::max(nsteps, last + 1);
        for (int64_t e = src_ptr[s]; e < src_ptr[s + 1]; ++e) need[P.pairs[e].tgt] = std::max(need[P.pairs[e].tgt], last + 1);
      }
    }
    P.hlevel_first_step[nLevels] = nsteps;
  }

  lap("pairs + schedule + steps");
  // ---- ownership for the sharded factorization ----------------------------------------------------------
  P.owner.assign(supNo, opt.world > 1 ? -1 : 0);
  if (opt.world > 1) {
    const int first_top = first_top_level;
    std::vector<int32_t> par(supNo, -1), root(supNo, -1);
    for (int s = 0; s < supNo; ++s) {
      const SupInfo& I = P.sup[s];
      if (I.r > I.w) par[s] = col2Sup[lR[I.rowptr + I.w]];   // supernodal etree parent = owner of the first row below
    }
    std::vector<double> cost;          // per bottom subtree, in column order of the roots
    std::vector<int32_t> roots;
    for (int s = supNo - 1; s >= 0; --s) {
      if (hl[s] >= first_top) continue;
      const int p2 = par[s];
      root[s] = (p2 < 0 || hl[p2] >= first_top) ? s : root[p2];
    }
    // cost of a subtree = flops its owner executes: POTRF + TRSM of its supernodes and EVERY update they push
    // (into the subtree and, fan-in, into the top separators)
    std::vector<double> rcost(supNo, 0.0);
    for (int s = 0; s < supNo; ++s) {
      if (root[s] < 0) continue;
      const double w = P.sup[s].w, r = P.sup[s].r;
      double c = w * w * w / 3.0 + w * w * (r - w);
      for (int64_t e = src_ptr[s]; e < src_ptr[s + 1]; ++e) {
        const double nd1 = P.pairs[e].nd1, nd3 = P.pairs[e].m - P.pairs[e].nd1;
        c += nd1 * nd1 * w + 2.0 * nd3 * nd1 * w;
      }
      rcost[root[s]] += c;
    }
    double total = 0;
    for (int s = 0; s < supNo; ++s) if (root[s] == s) { roots.push_back(s); total += rcost[s]; }
    // longest-processing-time first: subtrees by decreasing cost, each to the least loaded rank (ties: lowest rank).
    // A contiguous deal keeps a rank's panels in fewer runs but balances badly when the subtrees are few and unequal
    // (3D 27-pt 160^3, 9 subtrees, 8 ranks: max/mean 1.96 contiguous, 1.44 LPT; 3D 7-pt 100^3, 82 subtrees: 1.16 both).
    std::vector<int32_t> rown(supNo, -1);
    (void)total;
    {
      std::vector<int32_t> byc(roots);
      std::stable_sort(byc.begin(), byc.end(), [&](int32_t a, int32_t b2) { return rcost[a] > rcost[b2]; });
      std::vector<double> load(opt.world, 0.0);
      for (int s : byc) {
        int best = 0;
        for (int q = 1; q < opt.world; ++q) if (load[q] < load[best]) best = q;
        rown[s] = best;
        load[best] += rcost[s];
      }
    }
    for (int s = 0; s < supNo; ++s) if (root[s] >= 0) P.owner[s] = rown[root[s]];
    for (const PairDesc& q : P.pairs)
      if (P.owner[q.src] < 0 && P.owner[q.tgt] >= 0) { P.error = "top supernode updates a bottom one"; return PARSY_CUDA_ERR_BAD_SCHEDULE; }
  }
  std::vector<int32_t> node_first(supNo + 1, 0);
  for (int s = 0; s < supNo; ++s) node_first[s + 1] = node_first[s] + nblk[s];
  const bool dist_top = opt.world > 1 && opt.phase == 2 && opt.top_distributed != 0;
  P.node_owner.assign(node_first[supNo], -1);
  P.first_top_step = nsteps;
  if (opt.world > 1) {
    int rr = 0;
    for (int s = 0; s < supNo; ++s) {
      if (P.owner[s] >= 0) continue;
      P.first_top_step = std::min(P.first_top_step, (int)step0[s]);
      for (int b2 = 0; b2 < nblk[s]; ++b2) P.node_owner[node_first[s] + b2] = (rr++ / std::max(1, opt.top_chunk)) % opt.world;
    }
  }
  auto sup_active = [&](int s) {
    if (opt.phase == 1) return P.owner[s] == opt.rank;
    if (opt.phase == 2) return P.owner[s] < 0;
    return true;
  };
  // fan-in: an update runs where its SOURCE lives — bottom sources on their owner (targets: the same subtree or the
  // top separators, accumulated locally and summed over the ranks afterwards), top sources in phase 2
  auto pair_active = [&](int src, int tgt) {
    (void)tgt;
    if (opt.phase == 1) return P.owner[src] == opt.rank;
    if (opt.phase == 2) return P.owner[src] < 0;
    return true;
  };

  // ---- update pairs ---------------------------------------------------------------------------------
  for (const PairDesc& q : P.pairs) {
    const double k = P.sup[q.src].w, nd1 = q.nd1, nd3 = q.m - q.nd1;
    P.flops_update += nd1 * nd1 * k + 2.0 * nd3 * nd1 * k;
  }
  // pairs are emitted grouped by source d ascending; the step of a pair is the last step of its source
  // (the whole width of d is applied in one task, K = width(d)).

  lap("ownership + flops");
  // ---- bucket everything by step ----------------------------------------------------------------------
  P.steps.assign(nsteps, Step());
  for (int H = 0; H < nLevels; ++H)
    for (int st = P.hlevel_first_step[H]; st < P.hlevel_first_step[H + 1]; ++st) P.steps[st].hlevel = H;
  if (opt.ignore_hlevels) for (auto& s : P.steps) s.hlevel = 0;

  auto is_small_pair = [&](int K, int N) { return K <= 32 && N <= 32; };
  auto use128 = [&](int M, int N) { return N > 64 && (int64_t)M * N >= 2 * 128 * 128; };
  auto upd_flops = [](const GemmTask& t) { return (double)t.N * t.N * t.K + 2.0 * (double)(t.M - t.N) * t.N * t.K; };

  // factor-side lists: small supernodes and block columns per step
  std::vector<int32_t> c_small(nsteps, 0), c_blk(nsteps, 0);
  for (int s = 0; s < supNo; ++s) {
    if (!sup_active(s)) continue;
    if (P.sup[s].flags) { c_small[step0[s]]++; continue; }
    for (int b2 = 0; b2 < nblk[s]; ++b2) c_blk[step0[s] + b2]++;
  }
  std::vector<int64_t> o_small(nsteps + 1, 0), o_blk(nsteps + 1, 0);
  for (int st = 0; st < nsteps; ++st) { o_small[st + 1] = o_small[st] + c_small[st]; o_blk[st + 1] = o_blk[st] + c_blk[st]; }
  P.small_list.resize(o_small[nsteps]);
  P.block_tasks.resize(o_blk[nsteps]);
  for (int st = 0; st < nsteps; ++st) {
    P.steps[st].small_sup = Range{(int32_t)o_small[st], (int32_t)o_small[st + 1]};
    P.steps[st].blocks = Range{(int32_t)o_blk[st], (int32_t)o_blk[st + 1]};
  }

  // GEMM-shaped tasks, generated with a sort key: step, then class (0 trsm, 1 tiles128, 2 tiles64, 3 small), then
  // group (0 = "A": the target is factored in the very next step, 1 = "R": the rest; the executor overlaps R with
  // the next step's POTRF/TRSM on a second stream)
  // Kept compact: a generated task is (step, class, group) plus a reference — the index of a real pair whose GemmTask
  // is written straight into its final place after the sort (the warp-FMA pairs, 99 % of the list on 2-D problems), or
  // ~index into `extra`, the tasks that exist in full already (TRSM, trailing updates, tiled and split pairs).
  struct Gen { int32_t step; int32_t ref; int8_t cls, grp; };
  BigVec<GemmTask> extra;
  auto task_of = [&](const Gen& g) -> GemmTask& { return extra[~g.ref]; };
  // distributed top: does an update task of (step, group) read a panel that another rank factors (= must it wait for
  // that step's broadcasts)?  Set by the emitters below through `src_remote`.
  bool src_remote = false;
  BigVec<Gen> gen;
  gen.reserve(P.pairs.size() + 3 * (size_t)o_blk[nsteps]);
  // group of an update by the step its target is factored at: 0 = the very next step ("A", on the chain), 1 = soon
  // ("R"), 2 = at least FAR_STEPS later ("far": distributed top only — runs on its own low-priority stream and is
  // awaited FAR_STEPS steps later, so bursts of separator-to-separator updates do not stall the chain)
  auto group_of = [&](int st, int tstep) { return tstep <= st + 1 ? 0 : ((dist_top && tstep >= st + FAR_STEPS) ? 2 : 1); };
  auto emit_update = [&](const GemmTask& t, int st, int tstep, bool real_pair) {
    const int grp = group_of(st, tstep);
    Gen g; g.ref = ~(int32_t)extra.size(); g.step = st; g.grp = (int8_t)grp;
    extra.push_back(t);
    if (src_remote) P.steps[st].upd_remote[grp] = 1;
    const double fl = upd_flops(t);
    if (real_pair && is_small_pair(t.K, t.N)) { g.cls = 3; P.class_flops[5] += fl; P.n_pairs_small++; }
    else if (use128(t.M, t.N)) { g.cls = 1; P.class_flops[3] += fl; if (real_pair) P.n_pairs_tiled++; }
    else { g.cls = 2; P.class_flops[4] += fl; if (real_pair) P.n_pairs_tiled++; }
    gen.push_back(g);
  };

  std::vector<int32_t> f_small(nsteps, 0), f_blk(nsteps, 0);
  int32_t slot = 0;
  // distributed top: the block columns this rank owns come first in every step's list (the factor launches take that
  // prefix; the sweeps and the inverse blocks need all of them), hence two passes
  for (int pass = 0; pass < (dist_top ? 2 : 1); ++pass)
  for (int s = 0; s < supNo; ++s) {
    const SupInfo& I = P.sup[s];
    if (!sup_active(s)) continue;
    if (I.flags) {
      if (pass) continue;
      P.small_list[o_small[step0[s]] + f_small[step0[s]]++] = s;
      P.class_flops[0] += (double)I.w * I.w * I.w / 3.0 + (double)I.w * I.w * (I.r - I.w);
      continue;
    }
    for (int b2 = 0; b2 < nblk[s]; ++b2) {
      const bool mine = !dist_top || P.node_owner[node_first[s] + b2] == opt.rank;
      if (dist_top && mine != (pass == 0)) continue;
      const int st = step0[s] + b2, j0 = b2 * NB, nb = std::min(NB, I.w - j0);
      Step& S = P.steps[st];
      BlockTask bt; memset(&bt, 0, sizeof(bt));
      bt.sup = s; bt.j0 = j0; bt.nb = nb; bt.slot = slot;
      P.block_tasks[o_blk[st] + f_blk[st]++] = bt;
      S.max_nb = std::max(S.max_nb, nb);
      ++slot;
      if (!mine) { P.invert_tasks.push_back(bt); continue; }
      S.blocks_owned++;
      P.class_flops[1] += (double)nb * nb * nb / 3.0;
      const int Mb = I.r - j0 - nb;
      if (Mb > 0) {
        GemmTask t; memset(&t, 0, sizeof(t));
        t.a_off = I.valptr + (int64_t)j0 * I.r + j0 + nb; t.b_off = (int64_t)bt.slot * NB_MAX * NB_MAX; t.c_off = t.a_off;
        t.rel_off = -1; t.lda = I.r; t.ldb = NB_MAX; t.ldc = I.r; t.M = Mb; t.N = nb; t.K = nb;
        t.flags = GF_OVERWRITE | GF_B_LINV;
        Gen g; g.ref = ~(int32_t)extra.size(); g.step = st; g.cls = 0; g.grp = 0;
        extra.push_back(t);
        gen.push_back(g);
        P.class_flops[2] += (double)Mb * nb * nb;
      }
    }
  }
  // Trailing updates inside a wide supernode.  Block columns are grouped in runs of TRAIL_KBLK.  The panel of block
  // column b is applied right away (K = one block column) only to the targets that are factored soon — b+1 ("A", on
  // the chain) and the rest of b's run plus the first block column of the next run; everything further right receives
  // the whole run at once when its last block column is finished (K = TRAIL_KBLK block columns): four times fewer
  // scatter epilogues and operand re-reads per flop than rank-128 updates, with the same dependencies.
  constexpr int TRAIL_KBLK = 4;
  auto emit_trailing = [&](int s, int kb0, int kb1, int tb0, int tb1, int st, int tstep) {
    const SupInfo& I = P.sup[s];
    const int k0 = kb0 * NB, k1 = std::min(I.w, kb1 * NB), c0 = tb0 * NB, c1 = std::min(I.w, tb1 * NB);
    if (k1 <= k0 || c1 <= c0) return;
    GemmTask t; memset(&t, 0, sizeof(t));
    t.a_off = I.valptr + (int64_t)k0 * I.r + c0; t.b_off = t.a_off;
    t.c_off = I.valptr + (int64_t)c0 * I.r + c0;
    t.rel_off = -1; t.lda = t.ldb = t.ldc = I.r; t.M = I.r - c0; t.N = c1 - c0; t.K = k1 - k0; t.flags = GF_LOWER;
    emit_update(t, st, tstep, false);
  };
  for (int s = 0; s < supNo; ++s) {
    const SupInfo& I = P.sup[s];
    if (!sup_active(s) || I.flags) continue;
    const int nbk = nblk[s];
    for (int b2 = 0; b2 < nbk; ++b2) {
      const int st = step0[s] + b2;
      const int g0 = (b2 / TRAIL_KBLK) * TRAIL_KBLK, g1 = std::min(nbk, g0 + TRAIL_KBLK);
      const int near_end = std::min(nbk, g1 + 1);          // targets b2+1 .. near_end-1 get this panel alone
      const bool run_done = b2 == g1 - 1 && g1 + 1 < nbk;  // targets g1+1 .. get the whole run now
      if (dist_top) {
        // one task per target block column, kept only if this rank owns that block column
        auto remote = [&](int kb0, int kb1) { for (int kb = kb0; kb < kb1; ++kb) if (P.node_owner[node_first[s] + kb] != opt.rank) return true; return false; };
        for (int b3 = b2 + 1; b3 < near_end; ++b3) {
          if (P.node_owner[node_first[s] + b3] != opt.rank) continue;
          src_remote = remote(b2, b2 + 1);
          emit_trailing(s, b2, b2 + 1, b3, b3 + 1, st, step0[s] + b3);
        }
        for (int b3 = g1 + 1; run_done && b3 < nbk; ++b3) {
          if (P.node_owner[node_first[s] + b3] != opt.rank) continue;
          src_remote = remote(g0, g1);
          emit_trailing(s, g0, g1, b3, b3 + 1, st, step0[s] + b3);
        }
        src_remote = false;
      } else {
        if (b2 + 1 < nbk) emit_trailing(s, b2, b2 + 1, b2 + 1, b2 + 2, st, st + 1);
        if (b2 + 2 < near_end) emit_trailing(s, b2, b2 + 1, b2 + 2, near_end, st, st + 2);
        if (run_done) emit_trailing(s, g0, g1, g1 + 1, nbk, st, st + 2);
      }
    }
  }
  P.n_slots = slot; P.n_block_cols = slot;

  // narrow supernodes first inside every step (the low-register kernel variant takes the leading part of the list)
  std::vector<int32_t> sort_tmp;
  for (int st = 0; st < nsteps; ++st) {
    Step& S = P.steps[st];
    // stable counting sort by width (1 .. SMALL_W)
    int32_t* L0 = P.small_list.data() + S.small_sup.begin;
    const int cnt = S.small_sup.size();
    int32_t at[SMALL_W + 2] = {0};
    for (int i = 0; i < cnt; ++i) at[P.sup[L0[i]].w + 1]++;
    for (int w2 = 0; w2 <= SMALL_W; ++w2) at[w2 + 1] += at[w2];
    S.small_narrow = at[SMALL_W_NARROW + 1];
    sort_tmp.assign(L0, L0 + cnt);
    for (int i = 0; i < cnt; ++i) L0[at[P.sup[sort_tmp[i]].w]++] = sort_tmp[i];
  }
  // ---- dataflow sweeps: nodes, tasks in dependency order, target lists, counters ------------------------------
  int64_t sweep_violations = -1;
  auto build_sweeps = [&]() {
    const double t_sw0 = tnow();
    auto sw_at = [&](const char* what) { if (timing_on) fprintf(stderr, "[plan] (sweep thread: %s at) %.1f ms\n", what, (tnow() - t_sw0) * 1e3); };
    P.n_nodes = node_first[supNo];
    // node of every column, as one table: the target scans below look up every stored row index once
    std::vector<int32_t> node_of_col(n);
    for (int s = 0; s < supNo; ++s) {
      const SupInfo& I = P.sup[s];
      const int32_t nf = node_first[s];
      if (I.flags) std::fill(node_of_col.begin() + I.col0, node_of_col.begin() + I.col0 + I.w, nf);
      else for (int j = 0; j < I.w; ++j) node_of_col[I.col0 + j] = nf + j / NB;
    }
    auto node_of_row = [&](int row) { return node_of_col[row]; };
    P.node_need.assign(P.n_nodes, 0);
    P.node_tiles.assign(P.n_nodes, 1);
    P.solve_ctas.reserve((size_t)supNo / 4 + 2 * P.block_tasks.size());
    std::vector<uint8_t> is_wide(supNo);
    for (int s = 0; s < supNo; ++s) is_wide[s] = P.sup[s].flags ? 0 : 1;

    // Target list of one task: the nodes whose unknowns its rows touch, in row order.  Offsets are relative to `out`
    // (one list per thread, joined below).
    //   block-column slices: one entry (and one unit of `need`) per node AND 64-row tile, so that a tile can be published
    //     as soon as its own atomics are out — a scan of the slice's row indices;
    //   narrow supernodes: one entry per node — read off the update pairs of the supernode (one pair = one maximal run
    //     of rows inside one target supernode = one node, unless the target is wide: then its rows are scanned for the
    //     block-column boundaries).
    // Runs of one update pair, as fn(end row, node): its rows lie inside ONE target supernode, so a narrow target is a
    // single run; inside a wide target the (sorted) rows are cut at the block-column boundaries by bisection.
    auto pair_walk = [&](const PairDesc& q, const SupInfo& I, auto&& fn) {
      if (!is_wide[q.tgt]) { fn(q.lb + q.nd1, node_first[q.tgt]); return; }
      const int* rows = lR + I.rowptr;
      const int c0 = P.sup[q.tgt].col0, nf = node_first[q.tgt];
      int pos = q.lb;
      const int end = q.lb + q.nd1;
      while (pos < end) {
        const int blk = (rows[pos] - c0) / NB;
        // a block column has NB columns and the rows are distinct: its last row is less than NB entries away
        const int lim = std::min(end, pos + NB);
        const int next = rows[lim - 1] < c0 + (blk + 1) * NB ? lim : (int)(std::lower_bound(rows + pos, rows + lim, c0 + (blk + 1) * NB) - rows);
        fn(next, nf + blk);
        pos = next;
      }
    };
    // Row runs of the wide supernodes: maximal runs of rows (own columns included) inside one node, as (end row, node),
    // from the block columns and the update pairs of the supernode; the slices of its block columns — together they cover the rows of the
    // supernode nblk/2 times over — then walk the runs instead of the rows.
    std::vector<int32_t> wide_list, wide_id(supNo, -1);
    for (int s = 0; s < supNo; ++s) if (is_wide[s]) { wide_id[s] = (int32_t)wide_list.size(); wide_list.push_back(s); }
    std::vector<int64_t> run_ptr(wide_list.size() + 1, 0);
    BigVec<int32_t> run_end, run_node;
    {
      const int parts = std::max(1, plan_threads() / 2);
      std::vector<BigVec<int32_t>> le(parts), ln(parts);
      std::vector<size_t> pb(parts, wide_list.size()), pe(parts, wide_list.size());
      par_ranges(wide_list.size(), parts, [&](size_t w0, size_t w1, int k) {
        pb[k] = w0; pe[k] = w1;
        for (size_t wi = w0; wi < w1; ++wi) {
          const int s = wide_list[wi];
          const SupInfo& I = P.sup[s];
          const size_t before = le[k].size();
          for (int b2 = 0; b2 < nblk[s]; ++b2) { le[k].push_back(std::min(I.w, (b2 + 1) * NB)); ln[k].push_back(node_first[s] + b2); }   // own columns
          for (int64_t e2 = src_ptr[s]; e2 < src_ptr[s + 1]; ++e2)
            pair_walk(P.pairs[e2], I, [&](int end_row, int nd) { le[k].push_back(end_row); ln[k].push_back(nd); });
          run_ptr[wi + 1] = (int64_t)(le[k].size() - before);
        }
      });
      for (size_t wi = 0; wi < wide_list.size(); ++wi) run_ptr[wi + 1] += run_ptr[wi];
      run_end.resize((size_t)run_ptr[wide_list.size()]); run_node.resize(run_end.size());
      for (int k = 0; k < parts; ++k) {
        if (le[k].empty()) continue;
        memcpy(run_end.data() + run_ptr[pb[k]], le[k].data(), le[k].size() * sizeof(int32_t));
        memcpy(run_node.data() + run_ptr[pb[k]], ln[k].data(), ln[k].size() * sizeof(int32_t));
      }
    }
    sw_at("row runs");
    auto slice_targets = [&](SolveTask& t, BigVec<int32_t>& out, int32_t* need_local) {
      t.tgt_begin = (int32_t)out.size();
      const int64_t r0 = run_ptr[wide_id[t.sup]], r1 = run_ptr[wide_id[t.sup] + 1];
      const int32_t* re = run_end.data() + r0;
      const int32_t* rn = run_node.data() + r0;
      int k = (int)(std::upper_bound(re, re + (r1 - r0), t.row0) - re);      // the run holding row0
      const int end = t.row0 + t.nrows;
      for (int a = t.row0, tile = 0; a < end; a += SOLVE_TILE_ROWS, ++tile) {
        if (tile > 0) t.tile_tgt[tile - 1] = (int32_t)out.size();
        const int b2 = std::min(a + SOLVE_TILE_ROWS, end);
        for (;; ++k) {                                                        // every run that meets [a, b2)
          out.push_back(rn[k]); need_local[rn[k]]++;
          if (re[k] >= b2) break;
        }
        if (re[k] == b2) ++k;
      }
      t.tgt_end = (int32_t)out.size();
      for (int q = std::max(0, (t.nrows + SOLVE_TILE_ROWS - 1) / SOLVE_TILE_ROWS - 1); q < 4; ++q) t.tile_tgt[q] = t.tgt_end;
    };
    auto narrow_targets = [&](SolveTask& t, const SupInfo& I, BigVec<int32_t>& out, int32_t* need_local) {
      t.tgt_begin = (int32_t)out.size();
      for (int64_t e2 = src_ptr[t.sup]; e2 < src_ptr[t.sup + 1]; ++e2)
        pair_walk(P.pairs[e2], I, [&](int, int nd) { out.push_back(nd); need_local[nd]++; });
      t.tgt_end = (int32_t)out.size();
      for (int k = 0; k < 4; ++k) t.tile_tgt[k] = t.tgt_end;
    };
    // Leaf region: narrow supernodes without a block-column supernode anywhere below them in the supernodal etree.
    // It is closed under descendants, so its tasks can run first, on the light narrow-only kernels (backward sweep:
    // last); everything else — block columns and the narrow supernodes above them — follows on the general kernels.
    // Both parts keep the step order, which is a topological order of the dependencies.
    std::vector<char> above_block(supNo, 0);
    for (int s = 0; s < supNo; ++s) {
      const SupInfo& I = P.sup[s];
      if (!I.flags) above_block[s] = 1;
      if (above_block[s] && I.r > I.w) above_block[col2Sup[lR[I.rowptr + I.w]]] = 1;   // etree parent: first row below the block
    }
    P.n_narrow_prefix_ctas = 0;
    sw_at("tables");
    // First the ORDER of the tasks and their grouping into CTAs, as 16-byte descriptors (serial, cheap); the tasks
    // themselves and their target lists are then written by ranges of that list on several threads.
    struct TaskSrc { int32_t id; int32_t row0, nrows; int32_t kind; };   // id: supernode (kind 0) or block task (kind 1, 3 = first slice)
    BigVec<TaskSrc> src;
    src.reserve((size_t)supNo + 2 * P.block_tasks.size());
    for (int pass = 0; pass < 2; ++pass)
    for (int st = 0; st < nsteps; ++st) {
      const Step& S = P.steps[st];
      const bool seen_block = pass == 1;
      // narrow supernodes: long panels get a CTA each (kind 2), the others go eight per CTA (kind 0)
      auto narrow_task = [&](int s) {
        const SupInfo& I = P.sup[s];
        src.push_back(TaskSrc{s, I.w, I.r - I.w, 0});
      };
      auto is_tall = [&](int s) {
        const SupInfo& I = P.sup[s];
        return I.r - I.w > SOLVE_TALL_ROWS && (int64_t)(I.r - I.w) * I.w >= SOLVE_TALL_WORK;
      };
      std::vector<int> shortlist;
      for (int i = S.small_sup.begin; i < S.small_sup.end; ++i) {
        const int s = P.small_list[i];
        if ((above_block[s] != 0) != (pass == 1)) continue;
        if (!is_tall(s)) { shortlist.push_back(s); continue; }
        SolveCta c; c.kind = 2; c.first = (int32_t)src.size(); c.count = 1; c.pad = 0;
        narrow_task(s);
        P.solve_ctas.push_back(c);
        if (!seen_block) P.n_narrow_prefix_ctas++;
      }
      for (size_t i0 = 0; i0 < shortlist.size(); i0 += 8) {
        SolveCta c; c.kind = 0; c.first = (int32_t)src.size(); c.count = (int32_t)std::min<size_t>(8, shortlist.size() - i0); c.pad = 0;
        for (int i = 0; i < c.count; ++i) narrow_task(shortlist[i0 + i]);
        P.solve_ctas.push_back(c);
        if (!seen_block) P.n_narrow_prefix_ctas++;
      }
      for (int i = S.blocks.begin; pass == 1 && i < S.blocks.end; ++i) {
        const BlockTask& b = P.block_tasks[i];
        const SupInfo& I = P.sup[b.sup];
        const int below = I.r - b.j0 - b.nb;
        // graded slices: the rows right below the diagonal block belong to the next block column — the next link of
        // the dependency chain — so they go into short tasks whose whole register tile is fetched ahead of the
        // wait (SOLVE_CRITICAL_TASKS x SOLVE_TILE_ROWS rows); everything further down has slack and is cut into
        // SOLVE_TASK_ROWS-row tasks that amortise the diagonal-block solve
        std::vector<int> cut;   // first row (relative to the rows below the block) of every slice, plus the end
        cut.push_back(0);
        for (int k = 0; k < SOLVE_CRITICAL_TASKS && cut.back() + SOLVE_TILE_ROWS < below; ++k) cut.push_back(cut.back() + SOLVE_TILE_ROWS);
        while (cut.back() + SOLVE_TASK_ROWS < below) cut.push_back(cut.back() + SOLVE_TASK_ROWS);
        cut.push_back(std::max(below, 0));
        const int ntile = (int)cut.size() - 1;
        const int nd = node_first[b.sup] + b.j0 / NB;
        P.node_tiles[nd] = ntile;
        for (int k = 0; k < ntile; ++k) {
          SolveCta c; c.kind = 1; c.first = (int32_t)src.size(); c.count = 1; c.pad = 0;
          src.push_back(TaskSrc{i, b.j0 + b.nb + cut[k], cut[k + 1] - cut[k], k == 0 ? 3 : 1});
          P.solve_ctas.push_back(c);
        }
      }
    }
    sw_at("task order");
    const size_t ntask = src.size();
    P.solve_tasks.resize(ntask);
    const int parts = std::max(1, plan_threads() / 2);
    std::vector<BigVec<int32_t>> tg(parts);
    std::vector<size_t> part_begin(parts + 1, ntask), part_end(parts + 1, ntask);
    std::vector<std::vector<int32_t>> need_part(parts);      // counters per thread (the top separators are named by
                                                             // thousands of tasks: a shared counter would bounce), summed below
    par_ranges(ntask, parts, [&](size_t t0, size_t t1, int part) {
      part_begin[part] = t0; part_end[part] = t1;
      BigVec<int32_t>& out = tg[part];
      out.reserve((t1 - t0) * 3);
      need_part[part].assign(P.n_nodes, 0);
      int32_t* need_local = need_part[part].data();
      for (size_t ti = t0; ti < t1; ++ti) {
        const TaskSrc& d = src[ti];
        SolveTask& t = P.solve_tasks[ti];
        memset(&t, 0, sizeof(t));
        if (d.kind == 0) {
          const SupInfo& I = P.sup[d.id];
          t.sup = d.id; t.node = node_first[d.id]; t.j0 = 0; t.nb = I.w; t.slot = -1; t.row0 = d.row0; t.nrows = d.nrows; t.first = 1;
          t.rowptr = I.rowptr; t.valptr = I.valptr; t.col0 = I.col0; t.r = I.r;
          narrow_targets(t, I, out, need_local);
        } else {
          const BlockTask& bk = P.block_tasks[d.id];
          const SupInfo& I = P.sup[bk.sup];
          t.sup = bk.sup; t.node = node_first[bk.sup] + bk.j0 / NB; t.j0 = bk.j0; t.nb = bk.nb; t.slot = bk.slot;
          t.row0 = d.row0; t.nrows = d.nrows; t.first = d.kind == 3;
          t.rowptr = I.rowptr; t.valptr = I.valptr; t.col0 = I.col0; t.r = I.r;
          slice_targets(t, out, need_local);
        }
      }
    });
    for (int k = 0; k < parts; ++k)
      if (!need_part[k].empty()) for (int v = 0; v < P.n_nodes; ++v) P.node_need[v] += need_part[k][v];
    size_t total = 0;
    std::vector<size_t> base(parts + 1, 0);
    for (int k = 0; k < parts; ++k) { base[k] = total; total += tg[k].size(); }
    P.solve_targets.resize(total);
    // join the lists, shift the tasks' offsets, and hand every task its node's input count (complete now)
    par_ranges((size_t)parts, parts, [&](size_t k0, size_t k1, int) {
      for (size_t k = k0; k < k1; ++k) {
        if (!tg[k].empty()) memcpy(P.solve_targets.data() + base[k], tg[k].data(), tg[k].size() * sizeof(int32_t));
        const int32_t off = (int32_t)base[k];
        for (size_t ti = part_begin[k]; ti < part_end[k]; ++ti) {
          SolveTask& t = P.solve_tasks[ti];
          t.tgt_begin += off; t.tgt_end += off;
          for (int q = 0; q < 4; ++q) t.tile_tgt[q] += off;
          t.need = P.node_need[t.node];
        }
      }
    });
    sw_at("targets");
    // the sweep kernels spin on counters: a task list that is not a topological order would hang the device
    sweep_violations = sweep_order_violations(P);
    sw_at("order check, end");
  };
  // the sweep plan only depends on the factor-side lists above: built on a second thread next to the update lists
  std::exception_ptr sweep_error;
  std::thread sweep_thread([&] { try { build_sweeps(); } catch (...) { sweep_error = std::current_exception(); } });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } sweep_joiner{sweep_thread};
  lap("factor-side lists + trailing");
  // real pairs (+ relative-index bookkeeping)
  P.rel_prefix.assign(P.pairs.size() + 1, 0);
  P.rel_pair_src.resize(P.pairs.size()); P.rel_pair_tgt.resize(P.pairs.size()); P.rel_pair_lb.resize(P.pairs.size());
  int64_t rel = 0;
  for (size_t pi = 0; pi < P.pairs.size(); ++pi) {
    const PairDesc& q = P.pairs[pi];
    const SupInfo& D = P.sup[q.src]; const SupInfo& T = P.sup[q.tgt];
    const int st = step0[q.src] + nblk[q.src] - 1;
    P.rel_prefix[pi] = rel; P.rel_pair_src[pi] = q.src; P.rel_pair_tgt[pi] = q.tgt; P.rel_pair_lb[pi] = q.lb;
    const int64_t rel_here = rel;
    rel += q.m;
    if (!pair_active(q.src, q.tgt)) continue;
    src_remote = false;
    if (!dist_top && is_small_pair(D.w, q.nd1)) {      // the common case: no GemmTask yet, see pair_task below
      Gen g; g.step = st; g.ref = (int32_t)pi; g.cls = 3; g.grp = (int8_t)group_of(st, step0[q.tgt]);
      const double k = D.w, nd1 = q.nd1, nd3 = q.m - q.nd1;
      P.class_flops[5] += nd1 * nd1 * k + 2.0 * nd3 * nd1 * k; P.n_pairs_small++;
      gen.push_back(g);
      continue;
    }
    GemmTask t; memset(&t, 0, sizeof(t));
    t.a_off = D.valptr + q.lb; t.b_off = t.a_off; t.c_off = T.valptr; t.rel_off = rel_here;
    t.lda = t.ldb = D.r; t.ldc = T.r; t.M = q.m; t.N = q.nd1; t.K = D.w; t.flags = GF_LOWER | GF_ATOMIC;
    if (!dist_top) { emit_update(t, st, step0[q.tgt], true); continue; }
    // Distributed top.  (i) The pair's columns are split by the target block column they fall into; a rank keeps what
    // it owns.  (ii) A wide source is applied in K-chunks of PAIR_KBLK block columns as they are finished instead of
    // in one K = width task after its last block column: the ancestors' updates then overlap the source's own chain
    // instead of arriving in one burst right before the next separator starts (sums via red.add, so K can be cut).
    constexpr int PAIR_KBLK = 4;
    const int nkb = D.flags ? 1 : nblk[q.src];
    for (int n0 = 0; n0 < q.nd1;) {
      auto blk_of = [&](int j) { return T.flags ? 0 : (lR[D.rowptr + q.lb + j] - T.col0) / NB; };
      const int tb = blk_of(n0);
      int n1 = n0 + 1;
      while (n1 < q.nd1 && blk_of(n1) == tb) ++n1;
      if (P.node_owner[node_first[q.tgt] + tb] == opt.rank) {
        for (int kb0 = 0; kb0 < nkb; kb0 += PAIR_KBLK) {
          const int kb1 = std::min(nkb, kb0 + PAIR_KBLK);
          const int k0 = kb0 * NB, k1 = D.flags ? D.w : std::min(D.w, kb1 * NB);
          const int stq = step0[q.src] + kb1 - 1;
          GemmTask u = t;
          u.a_off += n0 + (int64_t)k0 * D.r; u.b_off = u.a_off; u.rel_off += n0; u.M = q.m - n0; u.N = n1 - n0; u.K = k1 - k0;
          src_remote = false;
          for (int kb = kb0; kb < kb1; ++kb) if (P.node_owner[node_first[q.src] + kb] != opt.rank) src_remote = true;
          emit_update(u, stq, step0[q.tgt] + tb, true);
        }
      }
      n0 = n1;
    }
  }
  src_remote = false;
  P.rel_prefix[P.pairs.size()] = rel;
  P.rel_entries = rel;

  lap("pair tasks");
  // tile size per launch: a (step, group) launch whose 128x64 / 64x64 tiles cannot cover the SMs is latency-bound by
  // the duration of one tile, so it is re-cut into 64x64 or 32x32 tiles (the panel -> next block column updates on the
  // critical path of a separator are the typical case)
  {
    constexpr int SMS = 148;
    std::vector<int32_t> tdef(3 * (size_t)nsteps, 0), t64(3 * (size_t)nsteps, 0);
    for (const Gen& g : gen) {
      if (g.cls != 1 && g.cls != 2) continue;
      const GemmTask& t = task_of(g);
      tdef[3 * g.step + g.grp] += g.cls == 1 ? lower_tiles(t.M, t.N, 128, 64) : lower_tiles(t.M, t.N, 64, 64);
      t64[3 * g.step + g.grp] += lower_tiles(t.M, t.N, 64, 64);
    }
    for (Gen& g : gen) {
      if (g.cls != 1 && g.cls != 2) continue;
      const int k = 3 * g.step + g.grp;
      if (tdef[k] >= SMS) continue;
      const double fl = upd_flops(task_of(g));
      const int ncls = t64[k] >= SMS ? 2 : 4;
      if (ncls == g.cls) continue;
      if (g.cls == 1) { P.class_flops[3] -= fl; P.class_flops[4] += fl; }
      g.cls = (int8_t)ncls;
    }
  }
  // split-K: a launch with fewer tiles than the GPU has CTA slots cannot fill the machine, and its tiles with a long
  // K (wide descendants) set the duration; the epilogue is an atomic add, so K can be cut into independent pieces
  {
    constexpr int FILL = 2 * 148;
    std::vector<int32_t> tl(3 * (size_t)nsteps, 0);
    auto ntiles = [&](const Gen& g) {
      const GemmTask& t = task_of(g);
      return g.cls == 1 ? lower_tiles(t.M, t.N, 128, 64) : g.cls == 2 ? lower_tiles(t.M, t.N, 64, 64) : lower_tiles(t.M, t.N, 32, 32);
    };
    for (const Gen& g : gen) if (g.cls == 1 || g.cls == 2 || g.cls == 4) tl[3 * g.step + g.grp] += ntiles(g);
    const size_t n0 = gen.size();
    for (size_t i = 0; i < n0; ++i) {
      if ((gen[i].cls != 1 && gen[i].cls != 2 && gen[i].cls != 4) || task_of(gen[i]).K < 128) continue;
      const int total = tl[3 * gen[i].step + gen[i].grp];
      if (total >= FILL) continue;
      const int K = task_of(gen[i]).K;
      const int f = std::min(std::min(cdiv(K, 64), cdiv(FILL, std::max(total, 1))), 8);
      if (f <= 1) continue;
      const int chunk = cdiv(cdiv(K, f), 16) * 16;
      for (int k0 = chunk; k0 < K; k0 += chunk) {
        Gen g = gen[i];
        GemmTask t = task_of(gen[i]);
        t.a_off += (int64_t)k0 * t.lda; t.b_off += (int64_t)k0 * t.ldb; t.K = std::min(chunk, K - k0);
        g.ref = ~(int32_t)extra.size();
        extra.push_back(t);
        gen.push_back(g);
      }
      task_of(gen[i]).K = chunk;
    }
  }
  lap("tile classes + split-K");
  if (gen.size() > (size_t)INT32_MAX) { P.error = "task list too large"; return PARSY_CUDA_ERR_BAD_ARG; }
  // stable counting sort by (step, class, group): the key space is small (15 buckets per step)
  BigVec<int32_t> ord(gen.size());
  std::vector<int32_t> bucket_begin;
  {
    auto bucket = [&](const Gen& g) { return (size_t)g.step * 15 + (size_t)g.cls * 3 + (size_t)g.grp; };
    std::vector<int32_t> start((size_t)nsteps * 15 + 1, 0);
    for (const Gen& g : gen) start[bucket(g) + 1]++;
    for (size_t b2 = 0; b2 + 1 < start.size(); ++b2) start[b2 + 1] += start[b2];
    bucket_begin = start;
    for (size_t i = 0; i < gen.size(); ++i) ord[start[bucket(gen[i])]++] = (int32_t)i;
  }
  lap("counting sort");
  P.gemm_tasks.resize(gen.size());
  lap("resize");
  // the sorted list, written in place by index ranges (threads); the per-step ranges are the bucket boundaries
  BigVec<uint32_t> shape(ord.size());      // per task: M, bit 31 = K <= 4 (all the small-task pass below needs)
  par_ranges(ord.size(), plan_threads(), [&](size_t i0, size_t i1, int) {
    for (size_t i = i0; i < i1; ++i) {
      const Gen& g = gen[ord[i]];
      if (g.ref < 0) { const GemmTask& e = extra[~g.ref]; P.gemm_tasks[i] = e; shape[i] = (uint32_t)e.M | (e.K <= 4 ? 0x80000000u : 0u); continue; }
      const PairDesc& q = P.pairs[g.ref];              // a warp-FMA pair, described by the pair table alone
      const SupInfo& D = P.sup[q.src];
      GemmTask& t = P.gemm_tasks[i];
      t.a_off = D.valptr + q.lb; t.b_off = t.a_off; t.c_off = P.sup[q.tgt].valptr; t.rel_off = P.rel_prefix[g.ref];
      t.lda = t.ldb = D.r; t.ldc = P.sup[q.tgt].r; t.M = q.m; t.N = q.nd1; t.K = D.w; t.flags = GF_LOWER | GF_ATOMIC;
      t.tile0 = 0;
      shape[i] = (uint32_t)t.M | (t.K <= 4 ? 0x80000000u : 0u);
    }
  });
  for (int st = 0; st < nsteps; ++st) {
    Step& S = P.steps[st];
    auto take = [&](int cls, int grp) { const size_t b2 = (size_t)st * 15 + cls * 3 + grp; return Range{bucket_begin[b2], bucket_begin[b2 + 1]}; };
    S.trsm = take(0, 0);
    for (int g = 0; g < 3; ++g) S.upd[g].u128 = take(1, g);
    for (int g = 0; g < 3; ++g) S.upd[g].u64 = take(2, g);
    for (int g = 0; g < 3; ++g) S.upd[g].small_pairs = take(3, g);
    for (int g = 0; g < 3; ++g) S.upd[g].u32 = take(4, g);
  }
  BigVec<Gen>().swap(gen);
  lap("sort + take");
  // row chunks of the small pairs, narrow (K <= 4) first inside every (step, group)
  {
    size_t total = 0;
    for (int st = 0; st < nsteps; ++st)
      for (int g = 0; g < 3; ++g)
        for (int gi = P.steps[st].upd[g].small_pairs.begin; gi < P.steps[st].upd[g].small_pairs.end; ++gi) total += cdiv((int)(shape[gi] & 0x7fffffffu), 256);
    P.small_tasks.reserve(total);
  }
  for (int st = 0; st < nsteps; ++st)
    for (int g = 0; g < 3; ++g) {
      UpdGroup& U = P.steps[st].upd[g];
      U.small.begin = (int32_t)P.small_tasks.size();
      for (int pass = 0; pass < 2; ++pass) {
        for (int gi = U.small_pairs.begin; gi < U.small_pairs.end; ++gi) {
          const bool narrow = (shape[gi] & 0x80000000u) != 0;
          const int M = (int)(shape[gi] & 0x7fffffffu);
          if (narrow != (pass == 0)) continue;
          for (int r0 = 0; r0 < M; r0 += 256) {
            SmallTask stt; stt.pair = gi; stt.row0 = r0; stt.nrows = std::min(256, M - r0); stt.pad = 0;
            P.small_tasks.push_back(stt);
          }
        }
        if (pass == 0) U.small_narrow = (int32_t)P.small_tasks.size() - U.small.begin;
      }
      U.small.end = (int32_t)P.small_tasks.size();
    }
  lap("small tasks");

  // tile prefixes per launch segment
  for (int st = 0; st < nsteps; ++st) {
    Step& S = P.steps[st];
    int32_t acc = 0;
    // TRSM row-tile height: the tallest of 64 / 32 / 16 that still gives every SM a tile
    S.trsm_tm = 64;
    for (int tm : {64, 32, 16}) {
      acc = 0;
      for (int i = S.trsm.begin; i < S.trsm.end; ++i) acc += cdiv(P.gemm_tasks[i].M, tm);
      S.trsm_tm = tm;
      if (acc >= 148) break;
    }
    acc = 0;
    for (int i = S.trsm.begin; i < S.trsm.end; ++i) { P.gemm_tasks[i].tile0 = acc; acc += cdiv(P.gemm_tasks[i].M, S.trsm_tm); }
    S.trsm_tiles = acc;
    for (int g = 0; g < 3; ++g) {
      UpdGroup& U = S.upd[g];
      acc = 0;
      for (int i = U.u128.begin; i < U.u128.end; ++i) { GemmTask& t = P.gemm_tasks[i]; t.tile0 = acc; acc += lower_tiles(t.M, t.N, 128, 64); }
      U.tiles128 = acc; acc = 0;
      for (int i = U.u64.begin; i < U.u64.end; ++i) { GemmTask& t = P.gemm_tasks[i]; t.tile0 = acc; acc += lower_tiles(t.M, t.N, 64, 64); }
      U.tiles64 = acc; acc = 0;
      for (int i = U.u32.begin; i < U.u32.end; ++i) { GemmTask& t = P.gemm_tasks[i]; t.tile0 = acc; acc += lower_tiles(t.M, t.N, 32, 32); }
      U.tiles32 = acc;
    }
    acc = 0;
    for (int i = S.blocks.begin; i < S.blocks.end; ++i) {
      BlockTask& bk = P.block_tasks[i];
      bk.tile0 = acc; acc += std::max(1, cdiv(P.sup[bk.sup].r - bk.j0 - bk.nb, 64));
    }
    S.solve_tiles = acc;
  }
  // ---- distributed top: panels to broadcast before each step ----------------------------------------------------
  P.bcast_ptr.assign(nsteps + 1, 0);
  if (dist_top) {
    std::vector<std::vector<int64_t>> per(nsteps);
    std::vector<std::vector<int32_t>> shp(nsteps);
    for (int s = 0; s < supNo; ++s) {
      if (P.owner[s] >= 0) continue;
      const SupInfo& I = P.sup[s];
      for (int b2 = 0; b2 < nblk[s]; ++b2) {
        const int j0 = I.flags ? 0 : b2 * NB, nb = I.flags ? I.w : std::min(NB, I.w - j0);
        auto& v = per[step0[s] + b2];
        v.push_back(P.node_owner[node_first[s] + b2]);
        v.push_back(I.valptr + (int64_t)j0 * I.r);
        v.push_back(I.valptr + (int64_t)(j0 + nb) * I.r);
        auto& h = shp[step0[s] + b2];
        h.push_back(j0); h.push_back(I.r); h.push_back(nb);
      }
    }
    for (int st = 0; st < nsteps; ++st) {
      P.bcast.insert(P.bcast.end(), per[st].begin(), per[st].end());
      P.bcast_shape.insert(P.bcast_shape.end(), shp[st].begin(), shp[st].end());
      P.bcast_ptr[st + 1] = (int32_t)(P.bcast.size() / 3);
    }
  }
  lap("prefixes + bcast");
  if (sweep_thread.joinable()) sweep_thread.join();
  if (sweep_error) std::rethrow_exception(sweep_error);
  lap("sweep plan");
  // ---- what a factorization zeroes / assembles, what the ranks sum, which columns the sweeps solve ---------------
  {
    auto runs_of = [&](auto pred, std::vector<int64_t>& out) {
      int64_t b = -1, e = -1;
      for (int s = 0; s <= supNo; ++s) {
        const bool in = s < supNo && pred(s);
        if (in) {
          const SupInfo& I = P.sup[s];
          if (b >= 0 && I.valptr != e) { out.push_back(b); out.push_back(e); b = -1; }
          if (b < 0) b = I.valptr;
          e = I.valptr + (int64_t)I.w * I.r;
        } else if (b >= 0) { out.push_back(b); out.push_back(e); b = -1; }
      }
    };
    P.skip_assemble.assign(supNo, 0);
    if (opt.world <= 1 || opt.phase == 0) { P.zero_runs = {0, P.xsize}; }
    else {
      // phase 1 prepares the buffer for both phases: the owned subtrees and the whole top are zeroed; A is scattered
      // into the owned subtrees and, on rank 0 only, into the top (the sum over the ranks must count A_top once)
      runs_of([&](int s) { return P.owner[s] == opt.rank || P.owner[s] < 0; }, P.zero_runs);
      for (int s = 0; s < supNo; ++s) P.skip_assemble[s] = !(P.owner[s] == opt.rank || (P.owner[s] < 0 && opt.rank == 0));
      runs_of([&](int s) { return P.owner[s] < 0; }, P.top_runs);
    }
    int32_t cb = -1, ce = -1;
    for (int s = 0; s <= supNo; ++s) {
      const bool in = s < supNo && sup_active(s);
      if (in) { if (cb < 0) cb = P.sup[s].col0; ce = P.sup[s].col0 + P.sup[s].w; }
      else if (cb >= 0) { P.col_runs.push_back(cb); P.col_runs.push_back(ce); cb = -1; }
    }
  }
  lap("runs");
  if (sweep_violations != 0) { P.error = "internal error: sweep task order is not a topological order"; return PARSY_CUDA_ERR_BAD_SCHEDULE; }
  return PARSY_CUDA_OK;
}

namespace {
struct Fnv {
  uint64_t h = 1469598103934665603ull;
  void bytes(const void* p, size_t nbytes) {
    const unsigned char* c = (const unsigned char*)p;
    // eight bytes per multiply: as good for change detection as the byte-wise variant and ~8x faster on long lists
    size_t i = 0;
    for (; i + 8 <= nbytes; i += 8) { uint64_t w; memcpy(&w, c + i, 8); h = (h ^ w) * 1099511628211ull; }
    for (; i < nbytes; ++i) h = (h ^ c[i]) * 1099511628211ull;
  }
  template <class T> void pod(const T& v) { bytes(&v, sizeof(T)); }
  template <class V> void vec(const V& v) { pod((uint64_t)v.size()); if (!v.empty()) bytes(v.data(), v.size() * sizeof(typename V::value_type)); }
};
}  // namespace

uint64_t plan_digest(const Plan& P) {
  Fnv f;
  f.pod(P.n); f.pod(P.nsuper); f.pod(P.nlevels); f.pod(P.xsize); f.pod(P.ssize); f.pod(P.nb);
  f.vec(P.sup); f.vec(P.small_list); f.vec(P.block_tasks); f.vec(P.gemm_tasks); f.vec(P.small_tasks);
  f.pod((uint64_t)P.steps.size());
  for (const Step& S : P.steps) {      // field by field: the struct has padding
    f.pod(S.hlevel); f.pod(S.small_sup); f.pod(S.small_narrow); f.pod(S.blocks); f.pod(S.blocks_owned); f.pod(S.trsm);
    f.pod(S.trsm_tiles); f.pod(S.trsm_tm);
    for (int g = 0; g < 3; ++g) {
      const UpdGroup& U = S.upd[g];
      f.pod(U.u128); f.pod(U.tiles128); f.pod(U.u64); f.pod(U.tiles64); f.pod(U.u32); f.pod(U.tiles32);
      f.pod(U.small_pairs); f.pod(U.small); f.pod(U.small_narrow); f.pod(S.upd_remote[g]);
    }
    f.pod(S.solve_tiles); f.pod(S.max_nb);
  }
  f.vec(P.hlevel_first_step); f.vec(P.pairs); f.vec(P.rel_prefix); f.vec(P.rel_pair_src); f.vec(P.rel_pair_tgt);
  f.vec(P.rel_pair_lb); f.pod(P.rel_entries); f.vec(P.owner); f.vec(P.node_owner); f.vec(P.bcast_ptr); f.vec(P.bcast);
  f.vec(P.bcast_shape); f.vec(P.invert_tasks); f.vec(P.zero_runs); f.vec(P.top_runs); f.vec(P.col_runs);
  f.vec(P.skip_assemble); f.pod(P.first_top_step); f.pod(P.n_nodes); f.vec(P.solve_tasks);
  f.pod((uint64_t)P.solve_ctas.size());
  for (const SolveCta& c : P.solve_ctas) { f.pod(c.kind); f.pod(c.first); f.pod(c.count); }
  f.pod(P.n_narrow_prefix_ctas); f.vec(P.solve_targets); f.vec(P.node_need); f.vec(P.node_tiles); f.pod(P.n_slots);
  f.pod(P.n_pairs); f.pod(P.n_pairs_small); f.pod(P.n_pairs_tiled); f.pod(P.n_block_cols);
  // the flop / byte totals are reporting only (and sums of doubles, sensitive to the order of addition): not hashed
  return f.h;
}

// Deadlock-freedom of the dataflow sweeps, checked on the host.  CTAs are handed out in list order (backward sweep: in
// reverse), resident CTAs spin until their inputs are complete, so every producer must come strictly earlier in the
// order than its consumers (tasks inside one CTA run side by side and must not depend on each other):
//   forward   a task that adds into node v (v in its target list) precedes every task of v, and the number of target
//             entries naming v equals need[v];
//   backward  every task of a node named in a task's target list comes later in the list (= earlier in the reverse walk);
//   split     the leaf-region prefix only depends on itself (forward), nothing outside it depends on ... the prefix
//             being run last (backward) — both follow from the two conditions above.
int64_t sweep_order_violations(const Plan& P) {
  const int64_t nct = (int64_t)P.solve_ctas.size();
  std::vector<int32_t> first_cta(P.n_nodes, INT32_MAX), last_cta(P.n_nodes, -1), seen(P.n_nodes, 0);
  std::vector<int32_t> cta_of(P.solve_tasks.size(), -1);
  int64_t bad = 0;
  for (int64_t c = 0; c < nct; ++c) {
    const SolveCta& C = P.solve_ctas[c];
    if (C.kind != 0 && C.count != 1) ++bad;
    if (c < P.n_narrow_prefix_ctas && C.kind == 1) ++bad;          // the light kernels have no block path
    for (int k = 0; k < C.count; ++k) {
      const int64_t ti = (int64_t)C.first + k;
      if (ti < 0 || ti >= (int64_t)P.solve_tasks.size() || cta_of[ti] != -1) { ++bad; continue; }
      cta_of[ti] = (int32_t)c;
      const int nd = P.solve_tasks[ti].node;
      first_cta[nd] = std::min(first_cta[nd], (int32_t)c);
      last_cta[nd] = std::max(last_cta[nd], (int32_t)c);
    }
  }
  for (size_t ti = 0; ti < P.solve_tasks.size(); ++ti) {
    const SolveTask& t = P.solve_tasks[ti];
    if (cta_of[ti] < 0) { ++bad; continue; }
    if (t.need != P.node_need[t.node]) ++bad;
    int32_t prev = t.tgt_begin;
    for (int k = 0; k < 4; ++k) { if (t.tile_tgt[k] < prev || t.tile_tgt[k] > t.tgt_end) ++bad; prev = t.tile_tgt[k]; }
    if (t.tile_tgt[3] != t.tgt_end) ++bad;
    for (int32_t q = t.tgt_begin; q < t.tgt_end; ++q) {
      const int v = P.solve_targets[q];
      if (v < 0 || v >= P.n_nodes) { ++bad; continue; }
      ++seen[v];
      if (!(cta_of[ti] < first_cta[v])) ++bad;     // forward: producer strictly before every task of v;
                                                   // backward: every task of v strictly after this consumer
    }
  }
  for (int v = 0; v < P.n_nodes; ++v) if (seen[v] != P.node_need[v]) ++bad;
  return bad;
}

}  // namespace parsy
