// Load-balanced level coarsening on the dependence DAG of a general lower-triangular CSC matrix: the inspector of
// lsolveParH2 (triangularSolve/Triangular_CSC.h:76) for inputs that are not Cholesky factors.
//
// Restates getCoarseLevelSet_DAG_CSC03 (cholesky/InspectionDAG_03.h:14-298; call site
// examples/triangularTest_DAG_nonChordal.cpp:343-360) so that levelPtr / parPtr / partition come out bit for bit:
//   1. wavefront level sets of the column DAG (buildLevelSet_CSC);
//   2. level cuts every divRate levels starting at minLevelDist (common/TreeUtils.h:257-324, heightPartitioning_DAG_Trng);
//      the number of bins of a cut is the width of the level below it;
//   3. per cut: connected components of the sub-DAG between two cuts by depth-first searches from the nodes of its
//      first level (common/DFS.h:99-139), components that meet are merged into the one with the smallest id; the nodes of
//      a component are listed in the order of a queue-driven topological sweep (common/BFS.h:63-90);
//   4. more components than bins: worst-fit bin packing, heaviest first (common/TreeUtils.h:218-255).
// Behaviours of the reference that shape the output and are kept: the column graph includes the diagonal entry (every
// in-degree counts it, a node is released when its count drops to 1); component costs and bin loads are never reset
// between cuts; the component counter is decremented on a merge without renumbering; the unstable std::sort of the
// components by cost.  Where the reference would index out of range (a component id beyond the counter after such a
// merge) this function reports an error instead.
//
// parsy_dag_lbc_bcsc is the supernodal twin, getCoarseLevelSet_DAG_BCSC02 (cholesky/Inspection_DAG_02.h:15-233; call
// site cholesky/LSparsity.h:1412, analyze_DAG): the same coarsening on the DAG of the BLOCKS of a BCSC factor.  What
// differs, and is kept: the level cuts come from the tree-LBC rule (heightPartitioning, common/TreeUtils.h:327-413);
// an edge is counted once per ROW of the source block that falls into the target block; merged components are
// relabelled from the second clash on, "the first is the minimum" being assumed (Inspection_DAG_02.h:127-137); packing
// happens when there are more components than innerParts, into the cut's own number of bins (:209-215).
#include <algorithm>
#include <climits>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/parsy_inspector.h"

void parsy_inspector_set_error(const std::string& msg);

namespace {
using ivec = std::vector<int>;

// The DAG as adjacency lists with repetitions and a self entry: node v lists idx[ptr[v] .. ptr[v+1]).
//   in-degree counts the entries from ptr[v] + indeg_off[v] on, the topological sweep walks from ptr[v] + sweep_off[v] on
//   (column form: both 0 — the diagonal entry is a self edge that the sweep removes itself; block form: the last
//   diagonal row is the one self entry that is counted and never removed).
struct Dag {
  int n = 0;
  ivec ptr, idx, indeg_off, sweep_off;
};
enum Flavor { COLUMNS_03, BLOCKS_02 };

int coarsen(const Dag& G, int H, const ivec& lvlPtr, const ivec& lvlSet, const ivec& cut, const ivec& bins, Flavor flavor,
            int innerParts, const double* nodeCost, int* nLevelsOut, int* levelPtrOut, int* parPtrOut, int* partitionOut) {
  const int n = G.n;
  ivec levelOf(n), indeg(n, 0);
  for (int l = 0; l < H; ++l) for (int k = lvlPtr[l]; k < lvlPtr[l + 1]; ++k) levelOf[lvlSet[k]] = l;
  for (int v = 0; v < n; ++v) for (int p = G.ptr[v] + G.indeg_off[v]; p < G.ptr[v + 1]; ++p) indeg[G.idx[p]]++;
  const int ncuts = (int)cut.size() - 1;

  ivec mark(n, 0), comp(n, -1), stack((size_t)2 * n);
  std::vector<char> seen(n, 0), uniq(n, 0);
  std::vector<double> compCost(n, 0.0), binLoad(n, 0.0);
  int* xi = stack.data();
  int* pstack = xi + n;
  int nparts = 0, filled = 0;
  levelPtrOut[0] = 0;
  parPtrOut[0] = 0;
  auto cost = [&](int v) { return nodeCost ? nodeCost[v] : 1.0; };

  for (int l = 0; l < ncuts; ++l) {
    const int first = cut[l], below = first - 1, upper = cut[l + 1];
    auto set_bounds = [&](int value, bool release) {
      for (int k = lvlPtr[std::max(below, 0)]; k < lvlPtr[below + 1]; ++k) mark[lvlSet[k]] = value;
      for (int lev = upper; lev < H; ++lev)
        for (int k = lvlPtr[lev]; k < lvlPtr[lev + 1]; ++k) {
          mark[lvlSet[k]] = value;
          if (release) seen[lvlSet[k]] = 1;      // keeps the sweep below inside this cut
        }
    };
    set_bounds(1, false);
    int ncomp = 0;
    ivec clash;
    if (first < H)
    for (int k = lvlPtr[first]; k < lvlPtr[first + 1]; ++k) {
      const int leaf = lvlSet[k];
      if (mark[leaf]) continue;
      // depth-first search from `leaf`; finished nodes are stacked at xi[top..n)
      int top = n, head = 0;
      xi[0] = leaf;
      while (head >= 0) {
        const int j = xi[head];
        if (!mark[j]) { mark[j] = 1; pstack[head] = G.ptr[j]; }
        if (mark[j] == -1) clash.push_back(j);
        bool done = true;
        for (int p = pstack[head]; p < G.ptr[j + 1]; ++p) {
          const int i = G.idx[p];
          if (mark[i] == -1) clash.push_back(i);      // belongs to a component found earlier in this cut
          if (mark[i]) continue;
          pstack[head] = p;
          xi[++head] = i;
          done = false;
          break;
        }
        if (done) { --head; xi[--top] = j; }
      }
      // distinct components among the clashes, smallest id
      int target = INT_MAX;
      for (size_t q = 0; q < clash.size();) {
        const int c = comp[clash[q]];
        if (!uniq[c]) { uniq[c] = 1; target = std::min(target, c); ++q; }
        else clash.erase(clash.begin() + (long)q);
      }
      for (int v : clash) uniq[comp[v]] = 0;
      int id;
      if (!clash.empty()) {
        if (flavor == COLUMNS_03) {
          for (int v : clash) {
            const int other = comp[v];
            if (other == target) continue;
            --ncomp;
            for (int u = 0; u < n; ++u) if (comp[u] == other) comp[u] = target;
          }
        } else {
          for (size_t q = 1; q < clash.size(); ++q) {
            const int other = comp[clash[q]];
            for (int u = 0; u < n; ++u) if (comp[u] == other) comp[u] = target;
          }
          ncomp -= (int)clash.size() - 1;
        }
        clash.clear();
        id = target;
      } else {
        id = ncomp++;
      }
      for (int q = top; q < n; ++q) {
        const int v = xi[q];
        comp[v] = id;
        compCost[id] += cost(v);
        if (levelOf[v] != first) mark[v] = -1;
      }
    }
    set_bounds(0, true);
    // nodes of every component in the order of a queue-driven topological sweep from the cut's first level
    if (ncomp < 0) { parsy_inspector_set_error("component bookkeeping of the reference is undefined for this input"); return 3; }
    std::vector<ivec> lists((size_t)ncomp);
    if (first < H) {
      ivec queue;
      for (int k = lvlPtr[first]; k < lvlPtr[first + 1]; ++k) {
        queue.clear();
        queue.push_back(lvlSet[k]);
        for (size_t qh = 0; qh < queue.size(); ++qh) {
          const int v = queue[qh];
          seen[v] = 1;
          if (comp[v] < 0 || comp[v] >= ncomp) { parsy_inspector_set_error("component bookkeeping of the reference is undefined for this input"); return 3; }
          lists[comp[v]].push_back(v);
          for (int p = G.ptr[v] + G.sweep_off[v]; p < G.ptr[v + 1]; ++p) {
            const int i = G.idx[p];
            if (--indeg[i] == 1 && !seen[i]) queue.push_back(i);
          }
        }
      }
    }
    for (int lev = upper; lev < H; ++lev) for (int k = lvlPtr[lev]; k < lvlPtr[lev + 1]; ++k) seen[lvlSet[k]] = 0;
    // bins
    std::vector<ivec> merged;
    const int nb = bins[l];
    if ((int)lists.size() > (flavor == COLUMNS_03 ? nb : innerParts)) {
      struct Item { double cost; int idx; };
      std::vector<Item> items(lists.size());
      for (size_t i = 0; i < lists.size(); ++i) items[i] = Item{compCost[i], (int)i};
      std::sort(items.begin(), items.end(), [](Item a, Item b) { return a.cost > b.cost; });
      merged.assign((size_t)nb, ivec());
      for (const Item& it : items) {
        double mn = INT_MAX;
        int best = 0;
        for (int b = 0; b < nb; ++b) if (binLoad[b] < mn) { mn = binLoad[b]; best = b; }
        binLoad[best] += it.cost;
        merged[best].insert(merged[best].end(), lists[it.idx].begin(), lists[it.idx].end());
      }
    } else {
      merged.swap(lists);
    }
    levelPtrOut[l + 1] = levelPtrOut[l] + (int)merged.size();
    for (const ivec& b : merged) {
      for (int v : b) {
        if (filled >= n) { parsy_inspector_set_error("a node was scheduled twice"); return 3; }
        partitionOut[filled++] = v;
        comp[v] = nparts;
      }
      parPtrOut[++nparts] = filled;
    }
  }
  if (filled != n) { parsy_inspector_set_error("schedule does not cover every node"); return 3; }
  *nLevelsOut = ncuts;
  return 0;
}
}  // namespace

// cut rule of the tree LBC (restated in inspector.cpp next to its first user)
int parsy_height_partitioning(int nwaves, const std::vector<int>& wptr, int H, int innerParts, int minLevelDist, int divRate,
                              std::vector<int>& sizes, std::vector<int>& bounds);

extern "C" int parsy_dag_lbc_csc(int n, const int* Lp, const int* Li, int innerParts, int minLevelDist, int divRate,
                                 const double* nodeCost, int* nLevelsOut, int* levelPtrOut, int* parPtrOut,
                                 int* partitionOut) {
  if (n <= 0 || !Lp || !Li || !nLevelsOut || !levelPtrOut || !parPtrOut || !partitionOut) { parsy_inspector_set_error("NULL or empty argument"); return 2; }
  if (innerParts < 1 || divRate < 1) { parsy_inspector_set_error("innerParts and divRate must be positive"); return 2; }
  ivec lvlPtr((size_t)n + 1), lvlSet((size_t)n);
  const int H = parsy_build_level_set_csc(n, Lp, Li, lvlPtr.data(), lvlSet.data());
  if (H < 0) return 2;
  // level cuts every divRate levels from minLevelDist on; bins of a cut = width of the level below it
  ivec cut, bins;
  cut.push_back(0);
  if (H <= minLevelDist) {
    cut.push_back(H);
    bins.push_back(1);
  } else {
    auto width = [&](int level) { return lvlPtr[level + 1] - lvlPtr[level]; };
    int t = minLevelDist;
    if (t > 0 && t < H) {
      cut.push_back(t);
      bins.push_back(width(t - 1) / 2 > 1 ? width(t - 1) : 1);
    }
    t += divRate;
    while (t < H - 1) {
      if (t < 1) { parsy_inspector_set_error("minLevelDist / divRate put a level cut below level 1"); return 2; }
      bins.push_back(std::max(1, width(t - 1)));
      cut.push_back(t);
      t += divRate;
    }
    cut.push_back(H + 1);
    bins.push_back(1);
  }
  Dag G;
  G.n = n;
  G.ptr.assign(Lp, Lp + n + 1);
  G.idx.assign(Li, Li + Lp[n]);
  G.indeg_off.assign(n, 0);
  G.sweep_off.assign(n, 0);
  return coarsen(G, H, lvlPtr, lvlSet, cut, bins, COLUMNS_03, innerParts, nodeCost, nLevelsOut, levelPtrOut, parPtrOut, partitionOut);
}

extern "C" int parsy_dag_lbc_bcsc(int nblocks, const size_t* Li_ptr, const int* lR, const int* blk2col, const int* col2blk,
                                  int innerParts, int minLevelDist, int divRate, const double* nodeCost, int* nLevelsOut,
                                  int* levelPtrOut, int* parPtrOut, int* partitionOut) {
  if (nblocks <= 0 || !Li_ptr || !lR || !blk2col || !col2blk || !nLevelsOut || !levelPtrOut || !parPtrOut || !partitionOut) { parsy_inspector_set_error("NULL or empty argument"); return 2; }
  if (innerParts < 1 || divRate < 2) { parsy_inspector_set_error("innerParts must be positive and divRate at least 2"); return 2; }
  const int n = nblocks;
  Dag G;
  G.n = n;
  G.ptr.resize((size_t)n + 1);
  G.indeg_off.resize(n);
  G.sweep_off.resize(n);
  for (int b = 0; b < n; ++b) {
    const int w = blk2col[b + 1] - blk2col[b];
    if (w <= 0) { parsy_inspector_set_error("empty block"); return 2; }
    G.ptr[b] = (int)Li_ptr[blk2col[b]];
    G.indeg_off[b] = w - 1;
    G.sweep_off[b] = w;
  }
  G.ptr[n] = (int)Li_ptr[blk2col[n]];
  G.idx.resize((size_t)G.ptr[n]);
  for (int p = 0; p < G.ptr[n]; ++p) G.idx[p] = col2blk[lR[p]];
  // wavefront level sets of the block DAG (buildLevelSet_BCSC, triangularSolve/Inspection_Level.h:65-150): a block's
  // round is one more than the latest round among the blocks it depends on; blocks of a round in increasing order
  ivec lev(n, 0), lvlPtr((size_t)n + 1, 0), lvlSet((size_t)n);
  int H = 0;
  for (int b = 0; b < n; ++b) {
    H = std::max(H, lev[b] + 1);
    for (int p = G.ptr[b] + G.sweep_off[b]; p < G.ptr[b + 1]; ++p) {
      const int t = G.idx[p];
      if (t <= b || t >= n) { parsy_inspector_set_error("row structure is not block lower triangular"); return 2; }
      lev[t] = std::max(lev[t], lev[b] + 1);
    }
  }
  for (int b = 0; b < n; ++b) lvlPtr[lev[b] + 1]++;
  for (int l = 0; l < H; ++l) lvlPtr[l + 1] += lvlPtr[l];
  {
    ivec fill(lvlPtr.begin(), lvlPtr.begin() + H);
    for (int b = 0; b < n; ++b) lvlSet[fill[lev[b]]++] = b;
  }
  ivec wptr(lvlPtr.begin(), lvlPtr.begin() + H + 1), sizes, cut;
  parsy_height_partitioning(H, wptr, H, innerParts, minLevelDist, divRate, sizes, cut);
  return coarsen(G, H, lvlPtr, lvlSet, cut, sizes, BLOCKS_02, innerParts, nodeCost, nLevelsOut, levelPtrOut, parPtrOut, partitionOut);
}
