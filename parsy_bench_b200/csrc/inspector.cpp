// libparsy_inspector: host-side symbolic inspector, written from the specification in SURVEY.md Appendix F.
// Output arrays must equal the reference's analyze_p2 (cholesky/LSparsity.h:256-842) bit for bit; stages whose
// result is mathematically unique use whatever algorithm is convenient, stages with tie-breaking rules follow
// the rule stated next to them (with the reference location).
#include "../../include/parsy_inspector.h"
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

// METIS from the CUDA toolkit (libmetis_static.a): 64-bit idx_t, the library the reference build links too.
extern "C" {
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options, int64_t* perm,
                 int64_t* iperm);
int METIS_SetDefaultOptions(int64_t* options);
}

namespace {

thread_local std::string g_err;
using ivec = std::vector<int>;
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Csc {
  ivec p, i;
  std::vector<double> x;
  std::vector<int64_t> src;   // position of every entry in the caller's arrays
};

// F = tril(P A P') by columns with ascending rows, S = triu(P A P') = F' by columns.
// The within-column order of `upper_unsorted` follows the reference's permuted symmetric transpose
// (cholesky/Transpose.h:204-540: columns visited in new order, entries in stored order, appended to column
// max(i,j)); only ereach_sn's output ORDER depends on it.
void permute_lower(int n, const int* Ap, const int* Ai, const double* Ax, const int* perm, Csc& upper_unsorted,
                   Csc& lower_sorted) {
  ivec pinv(n);
  for (int k = 0; k < n; ++k) pinv[perm[k]] = k;
  const int64_t nnz = Ap[n];
  Csc& U = upper_unsorted;
  U.p.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    const int jold = perm[j];
    for (int q = Ap[jold]; q < Ap[jold + 1]; ++q)
      if (Ai[q] >= jold) U.p[std::max(pinv[Ai[q]], j) + 1]++;
  }
  for (int j = 0; j < n; ++j) U.p[j + 1] += U.p[j];
  U.i.resize(U.p[n]); U.x.resize(U.p[n]); U.src.resize(U.p[n]);
  ivec fill(U.p.begin(), U.p.end() - 1);
  for (int j = 0; j < n; ++j) {
    const int jold = perm[j];
    for (int q = Ap[jold]; q < Ap[jold + 1]; ++q) {
      if (Ai[q] < jold) continue;
      const int i = pinv[Ai[q]];
      const int dst = fill[std::max(i, j)]++;
      U.i[dst] = std::min(i, j);
      U.x[dst] = Ax ? Ax[q] : 0.0;
      U.src[dst] = q;
    }
  }
  (void)nnz;
  // transpose of the upper matrix, columns visited in order -> rows ascending
  Csc& L = lower_sorted;
  L.p.assign(n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int q = U.p[j]; q < U.p[j + 1]; ++q) L.p[U.i[q] + 1]++;
  for (int j = 0; j < n; ++j) L.p[j + 1] += L.p[j];
  L.i.resize(L.p[n]); L.x.resize(L.p[n]); L.src.resize(L.p[n]);
  ivec f2(L.p.begin(), L.p.end() - 1);
  for (int j = 0; j < n; ++j)
    for (int q = U.p[j]; q < U.p[j + 1]; ++q) {
      const int dst = f2[U.i[q]]++;
      L.i[dst] = j; L.x[dst] = U.x[q]; L.src[dst] = U.src[q];
    }
}

// Liu's elimination tree from the upper pattern (columns hold rows i <= k).  common/Etree.h:56
void etree_upper(int n, const Csc& U, ivec& parent) {
  parent.assign(n, -1);
  ivec anc(n, -1);
  for (int k = 0; k < n; ++k)
    for (int q = U.p[k]; q < U.p[k + 1]; ++q) {
      int i = U.i[q];
      while (i != -1 && i < k) {
        const int nxt = anc[i];
        anc[i] = k;
        if (nxt == -1) parent[i] = k;
        i = nxt;
      }
    }
}

// Postorder with children visited in (weight ascending, index ascending) order, or plain ascending index
// when weight == nullptr; roots in ascending index.  common/PostOrder.h:74-151
void postorder(int n, const ivec& parent, const int* weight, ivec& post) {
  ivec head(n, -1), next(n, -1);
  if (!weight) {
    for (int j = n - 1; j >= 0; --j)
      if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
  } else {
    // stable counting sort by clamped weight, heaviest first, so that pushing to the list heads leaves the
    // lightest child (smallest index among equals) first
    ivec bucket(n + 1, 0);
    auto wt = [&](int j) { return std::min(std::max(weight[j], 0), n - 1); };
    for (int j = 0; j < n; ++j) if (parent[j] >= 0) bucket[wt(j) + 1]++;
    for (int w = 0; w < n; ++w) bucket[w + 1] += bucket[w];
    ivec order(bucket[n]);
    {
      ivec f(bucket.begin(), bucket.end() - 1);
      for (int j = 0; j < n; ++j) if (parent[j] >= 0) order[f[wt(j)]++] = j;   // (weight asc, index asc)
    }
    for (int t = (int)order.size() - 1; t >= 0; --t) {
      const int j = order[t];
      next[j] = head[parent[j]]; head[parent[j]] = j;
    }
  }
  post.resize(n);
  ivec stack; stack.reserve(64);
  int k = 0;
  for (int r = 0; r < n; ++r) {
    if (parent[r] != -1) continue;
    stack.push_back(r);
    while (!stack.empty()) {
      const int p = stack.back();
      const int c = head[p];
      if (c == -1) { stack.pop_back(); post[k++] = p; }
      else { head[p] = next[c]; stack.push_back(c); }
    }
  }
}

// Column counts of L (diagonal included) from the lower pattern, etree and a postorder: skeleton-matrix
// algorithm of Gilbert, Ng & Peyton.  Result is unique.  cholesky/ColumnCount.h:141
void column_counts(int n, const Csc& Lw, const ivec& parent, const ivec& post, ivec& cc) {
  ivec first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n), delta(n, 0);
  for (int k = 0; k < n; ++k) {
    int j = post[k];
    delta[j] = (first[j] == -1) ? 1 : 0;
    for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
  }
  for (int i = 0; i < n; ++i) anc[i] = i;
  for (int k = 0; k < n; ++k) {
    const int j = post[k];
    if (parent[j] != -1) delta[parent[j]]--;
    for (int q = Lw.p[j]; q < Lw.p[j + 1]; ++q) {
      const int i = Lw.i[q];
      if (i <= j || first[j] <= maxfirst[i]) continue;   // j is not a leaf of the i-th row subtree
      maxfirst[i] = first[j];
      const int jprev = prevleaf[i];
      prevleaf[i] = j;
      delta[j]++;
      if (jprev != -1) {
        int qq = jprev;
        while (qq != anc[qq]) qq = anc[qq];
        for (int s = jprev; s != qq;) { const int sp = anc[s]; anc[s] = qq; s = sp; }
        delta[qq]--;
      }
    }
    if (parent[j] != -1) anc[j] = parent[j];
  }
  cc = delta;
  // children precede parents in `post`
  for (int k = 0; k < n; ++k) { const int j = post[k]; if (parent[j] != -1) cc[parent[j]] += cc[j]; }
}

struct Supernodes {
  ivec super, snz, sparent, supermap;
  int nsuper = 0;
};

// Fundamental supernodes + relaxed amalgamation.  cholesky/Inspection_BlockC.h:296-540
void find_supernodes(int n, const ivec& parent, const ivec& cc, const int nrelax[3], const double zrelax[3],
                     Supernodes& out) {
  ivec nchild(n, 0);
  for (int j = 0; j < n; ++j) if (parent[j] != -1) nchild[parent[j]]++;
  ivec fsuper;
  if (n > 0) fsuper.push_back(0);
  for (int j = 1; j < n; ++j)
    if (parent[j - 1] != j || cc[j - 1] != cc[j] + 1 || nchild[j] > 1) fsuper.push_back(j);
  const int nf = (int)fsuper.size();
  fsuper.push_back(n);
  ivec smap(n);
  for (int s = 0; s < nf; ++s) for (int k = fsuper[s]; k < fsuper[s + 1]; ++k) smap[k] = s;
  ivec sparent(nf), merged(nf, -1), nscol(nf), zeros(nf, 0), snz(nf);
  for (int s = 0; s < nf; ++s) {
    const int pj = parent[fsuper[s + 1] - 1];
    sparent[s] = pj == -1 ? -1 : smap[pj];
    nscol[s] = fsuper[s + 1] - fsuper[s];
    snz[s] = cc[fsuper[s]];
  }
  for (int s = nf - 2; s >= 0; --s) {
    if (sparent[s] == -1) continue;
    int cur = sparent[s];
    while (merged[cur] != -1) cur = merged[cur];
    for (int t = sparent[s]; merged[t] != -1;) { const int nx = merged[t]; merged[t] = cur; t = nx; }
    if (cur != s + 1) continue;
    const int ns0 = nscol[s], ns1 = nscol[s + 1], ns = ns0 + ns1;
    int totzeros = zeros[s + 1];
    const double lnz1 = (double)snz[s + 1];
    bool merge;
    if (ns <= nrelax[0]) {
      merge = true;   // tiny supernodes always merge; the zero count is NOT increased on this path (:424-427)
    } else {
      const double lnz0 = snz[s];
      const double xnewzeros = ns0 * (lnz1 + ns0 - lnz0);
      const int newzeros = ns0 * (snz[s + 1] + ns0 - snz[s]);
      if (xnewzeros == 0) {
        merge = true;
      } else {
        const double xtotzeros = (double)totzeros + xnewzeros;
        const double xns = (double)ns;
        const double xtotsize = (xns * (xns + 1) / 2) + xns * (lnz1 - ns1);
        const double z = xtotzeros / xtotsize;
        totzeros += newzeros;
        merge = ((ns <= nrelax[1] && z < zrelax[0]) || (ns <= nrelax[2] && z < zrelax[1]) || (z < zrelax[2])) &&
                (xtotsize < (double)(INT_MAX / sizeof(double)));
      }
    }
    if (merge) {
      zeros[s] = totzeros;
      merged[s + 1] = s;
      snz[s] = ns0 + snz[s + 1];
      nscol[s] += nscol[s + 1];
    }
  }
  out.super.clear(); out.snz.clear();
  for (int s = 0; s < nf; ++s)
    if (merged[s] == -1) { out.super.push_back(fsuper[s]); out.snz.push_back(snz[s]); }
  out.nsuper = (int)out.super.size();
  out.super.push_back(n);
  out.supermap.resize(n);
  for (int s = 0; s < out.nsuper; ++s) for (int k = out.super[s]; k < out.super[s + 1]; ++k) out.supermap[k] = s;
  out.sparent.resize(out.nsuper);
  for (int s = 0; s < out.nsuper; ++s) {
    const int pj = parent[out.super[s + 1] - 1];
    out.sparent[s] = pj == -1 ? -1 : out.supermap[pj];
  }
}

// ---- LBC: Load-Balanced level Coarsening (cholesky/InspectionLevel_06.h:18-391) ------------------------------

// bottom-up waves of a forest: wave 0 = childless nodes in ascending index, ...  common/TreeUtils.h:119-169
int tree_waves(int n, const int* tree, ivec& wptr, ivec& wset, ivec& node2wave) {
  ivec nchild(n, 0);
  for (int k = 0; k < n; ++k) if (tree[k] >= 0) nchild[tree[k]]++;
  node2wave.assign(n, 0);
  // height above the leaves; parents have larger indices than children in an etree, but do not rely on it
  ivec ready;
  wset.clear(); wset.reserve(n);
  wptr.assign(1, 0);
  std::vector<char> done(n, 0);
  ivec cur;
  for (int i = 0; i < n; ++i) if (nchild[i] == 0) cur.push_back(i);
  int wave = 0;
  while (!cur.empty()) {
    std::sort(cur.begin(), cur.end());
    ivec nxt;
    for (int v : cur) { wset.push_back(v); node2wave[v] = wave; }
    for (int v : cur) if (tree[v] >= 0 && --nchild[tree[v]] == 0) nxt.push_back(tree[v]);
    wptr.push_back((int)wset.size());
    cur.swap(nxt);
    ++wave;
  }
  return wave;
}

// Cuts the wave range into clusters.  common/TreeUtils.h:327-413 (quirk B.6: one size entry per halving step)
int height_partitioning(int nwaves, const ivec& wptr, int H, int innerParts, int minLevelDist, int divRate,
                        ivec& sizes, ivec& bounds) {
  bounds.assign(1, 0);
  sizes.clear();
  if (nwaves <= 2) { bounds.push_back(nwaves); sizes.push_back(1); return 1; }
  auto cnt = [&](int w) -> long { return (w >= 0 && w < nwaves) ? (long)(wptr[w + 1] - wptr[w]) : LONG_MAX; };
  int ip = innerParts;
  while (ip > 1) {
    int cut = 0;
    // an index outside the wave array ends the scan (the reference reads past the array there, App. F)
    while (cnt(H - cut - 1) <= ip && cut < nwaves) ++cut;
    sizes.push_back(ip);
    const int t = H - cut - minLevelDist;
    if (t > bounds.back() && t < nwaves) bounds.push_back(t);
    ip /= divRate;
  }
  bounds.push_back(H + 1);
  sizes.push_back(1);
  return (int)bounds.size() - 1;
}

// Splits the cluster forest into connected subtrees; emission order = execution order inside a w-partition.
// State machine of cholesky/PostOrderSpliting.h:36-110 (front-inserting work list, leaves first).
void split_forest(int n, const ivec& tree, const std::vector<double>& cost, const ivec& cptr, const ivec& cidx,
                  ivec& nchild, std::vector<ivec>& parts, std::vector<double>& partCost) {
  std::vector<char> visited(n, 0), queued(n, 0);
  std::deque<int> work;
  parts.clear(); partCost.clear();
  for (int start = 0; start < n; ++start) {
    if (tree[start] == -2 || visited[start]) continue;
    parts.emplace_back(); partCost.push_back(0.0);
    ivec& part = parts.back();
    double& pc = partCost.back();
    work.push_back(start);
    while (!work.empty()) {
      const int k = work.front();
      if (nchild[k] == 0) {
        work.pop_front();
        part.push_back(k); pc += cost[k]; visited[k] = 1;
        const int par = tree[k];
        if (par < 0) break;             // root of this subtree: the part is complete
        nchild[par]--;
        if (!visited[par]) {
          if (!queued[par]) work.push_front(par);
          queued[par] = 1;
        }
      } else {
        for (int q = cptr[k]; q < cptr[k + 1]; ++q) {
          const int c = cidx[q];
          if (visited[c]) continue;
          if (nchild[c] == 0) {
            if (!visited[k]) {
              int t = k;
              if (!work.empty())
                while (t != work.back()) { queued[t] = 1; t = tree[t]; }
              queued[t] = 1;
            }
            part.push_back(c); pc += cost[c]; visited[c] = 1;
            nchild[k]--;
          } else {
            work.push_front(c);
          }
        }
      }
    }
  }
}

struct Schedule { ivec levelPtr, parPtr, partition; };

void lbc_schedule(int n, const ivec& etree, const ivec& super, const std::vector<double>& nodeCost, int innerParts,
                  int minLevelDist, int divRate, Schedule& out) {
  (void)super;
  ivec depth(n, 0);
  int H = 0;
  {
    // depth below the root in edges; H = deepest childless node (getTreeHeight, TreeUtils.h:87)
    ivec nchild(n, 0);
    for (int i = 0; i < n; ++i) if (etree[i] >= 0) nchild[etree[i]]++;
    for (int i = n - 1; i >= 0; --i) depth[i] = etree[i] >= 0 ? depth[etree[i]] + 1 : 0;   // parent index > child
    for (int i = 0; i < n; ++i) if (nchild[i] == 0) H = std::max(H, depth[i]);
  }
  ivec wptr, wset, node2wave;
  const int nwaves = tree_waves(n, etree.data(), wptr, wset, node2wave);
  ivec sizes, bounds;
  const int nclusters = height_partitioning(nwaves, wptr, H, innerParts, minLevelDist, divRate, sizes, bounds);

  out.levelPtr.assign(1, 0); out.parPtr.assign(1, 0); out.partition.clear(); out.partition.reserve(n);
  std::vector<double> binLoad(n, 0.0);   // carried over from cluster to cluster (quirk B.4)
  ivec tmpTree(n, -2), nchild(n), cptr(n + 1), cidx(n);
  for (int l = 0; l < nclusters; ++l) {
    const int bins = sizes[l];
    const int lo = bounds[l], hi = bounds[l + 1];
    std::fill(tmpTree.begin(), tmpTree.end(), -2);
    for (int w = lo; w < hi && w < nwaves; ++w)
      for (int q = wptr[w]; q < wptr[w + 1]; ++q) {
        const int v = wset[q], par = etree[v];
        const int pw = par >= 0 ? node2wave[par] : H + 1;
        tmpTree[v] = (pw < hi && pw >= lo) ? par : -1;     // edges leaving the cluster are cut
      }
    std::fill(nchild.begin(), nchild.end(), 0);
    for (int k = 0; k < n; ++k) if (tmpTree[k] >= 0) nchild[tmpTree[k]]++;
    cptr[0] = 0;
    for (int k = 0; k < n; ++k) cptr[k + 1] = cptr[k] + nchild[k];
    {
      ivec f(cptr.begin(), cptr.end() - 1);
      for (int k = 0; k < n; ++k) if (tmpTree[k] >= 0) cidx[f[tmpTree[k]]++] = k;   // children ascending
    }
    std::vector<ivec> parts;
    std::vector<double> partCost;
    split_forest(n, tmpTree, nodeCost, cptr, cidx, nchild, parts, partCost);
    std::vector<ivec> merged;
    if ((int)parts.size() > bins) {
      // worst-fit: heaviest subtree first into the currently lightest bin (first minimum, seeded with INT_MAX);
      // std::sort is not stable — equal costs land in libstdc++'s introsort order, as in the reference
      // (common/TreeUtils.h:205-255)
      struct Item { double cost; int idx; };
      std::vector<Item> items(parts.size());
      for (size_t i = 0; i < parts.size(); ++i) items[i] = Item{partCost[i], (int)i};
      std::sort(items.begin(), items.end(), [](Item a, Item b) { return a.cost > b.cost; });
      merged.assign(bins, ivec());
      for (const Item& it : items) {
        double mn = INT_MAX;
        int best = 0;
        for (int b = 0; b < bins; ++b) if (binLoad[b] < mn) { mn = binLoad[b]; best = b; }
        binLoad[best] += it.cost;
        merged[best].insert(merged[best].end(), parts[it.idx].begin(), parts[it.idx].end());
      }
    } else {
      merged.swap(parts);
    }
    out.levelPtr.push_back(out.levelPtr.back() + (int)merged.size());
    for (const ivec& b : merged) {
      out.partition.insert(out.partition.end(), b.begin(), b.end());
      out.parPtr.push_back((int)out.partition.size());
    }
  }
}

}  // namespace
// shared with dag_lbc.cpp (the block DAG coarsening cuts its levels with the same rule)
int parsy_height_partitioning(int nwaves, const std::vector<int>& wptr, int H, int innerParts, int minLevelDist, int divRate,
                              std::vector<int>& sizes, std::vector<int>& bounds) {
  return height_partitioning(nwaves, wptr, H, innerParts, minLevelDist, divRate, sizes, bounds);
}
namespace {

template <class T> T* dup(const std::vector<T>& v) {
  T* p = new T[std::max<size_t>(v.size(), 1)];
  if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

}  // namespace

// Level sets of the dependence DAG of a lower-triangular CSC matrix (diagonal first in every column), what
// buildLevelSet_CSC (triangularSolve/Inspection_Level.h:12-59) hands to lsolvePar (examples/triangularTest_DAG.cpp:171-175).
// The reference peels wavefronts: every round takes, in increasing index order, all columns with no unprocessed
// off-diagonal entry in their row, then removes their edges.  A column's round is therefore 1 + the latest round among
// the columns it depends on (longest path from a source), which is what is computed here in one pass over the columns —
// same levelPtr / levelSet, no repeated scans of the index range.
extern "C" int parsy_build_level_set_csc(int n, const int* Lp, const int* Li, int* levelPtr, int* levelSet) {
  if (n < 0 || !Lp || !Li || !levelPtr || !levelSet) { g_err = "NULL argument"; return -1; }
  std::vector<int> lev(n, 0);
  int levels = 0;
  for (int j = 0; j < n; ++j) {
    if (Lp[j + 1] <= Lp[j] || Li[Lp[j]] != j) { g_err = "every column must start with its diagonal entry"; return -1; }
    const int mine = lev[j];
    levels = std::max(levels, mine + 1);
    for (int p = Lp[j] + 1; p < Lp[j + 1]; ++p) {
      const int i = Li[p];
      if (i <= j || i >= n) { g_err = "matrix is not lower triangular"; return -1; }
      lev[i] = std::max(lev[i], mine + 1);
    }
  }
  std::fill(levelPtr, levelPtr + n + 1, 0);
  for (int j = 0; j < n; ++j) levelPtr[lev[j] + 1]++;
  for (int l = 0; l < levels; ++l) levelPtr[l + 1] += levelPtr[l];
  std::vector<int> fill(levelPtr, levelPtr + levels);
  for (int j = 0; j < n; ++j) levelSet[fill[lev[j]]++] = j;
  return levels;
}

extern "C" const char* parsy_inspector_last_error(void) { return g_err.c_str(); }
void parsy_inspector_set_error(const std::string& msg) { g_err = msg; }   // used by mmio.cpp

extern "C" void parsy_symbolic_free(parsy_symbolic* s) {
  if (!s) return;
  delete[] s->Perm; delete[] s->ColCount; delete[] s->Parent; delete[] s->super; delete[] s->sParent;
  delete[] s->col2Sup; delete[] s->pi; delete[] s->s; delete[] s->p; delete[] s->i_ptr; delete[] s->levelPtr;
  delete[] s->parPtr; delete[] s->partition; delete[] s->A1_p; delete[] s->A1_i; delete[] s->A2_p; delete[] s->A2_i;
  delete[] s->A2_x; delete[] s->A2_src;
  delete s;
}

extern "C" int parsy_inspect(int n, const int* Ap, const int* Ai, const double* Ax, int costParam, int levelParam,
                             int divRate, const int* userPerm, parsy_symbolic** out) {
  if (!out) { g_err = "out is NULL"; return 2; }
  *out = nullptr;
  if (n <= 0 || !Ap || !Ai) { g_err = "empty or NULL matrix"; return 2; }
  if (divRate < 2) { g_err = "divRate must be >= 2 (the level cut never terminates otherwise)"; return 2; }
  const double t0 = now();
  const int64_t nnz = Ap[n];
  for (int j = 0; j < n; ++j)
    if (Ap[j + 1] <= Ap[j] || Ai[Ap[j]] != j) { g_err = "every column must start with its diagonal entry"; return 2; }

  // ---- ordering: METIS_NodeND on the diagonal-free full graph (LSparsity.h:559-606) ----------------------
  ivec perm(n);
  double t_ord = 0;
  if (userPerm) {
    std::vector<char> seen(n, 0);
    for (int k = 0; k < n; ++k) {
      if (userPerm[k] < 0 || userPerm[k] >= n || seen[userPerm[k]]) { g_err = "userPerm is not a permutation"; return 2; }
      seen[userPerm[k]] = 1; perm[k] = userPerm[k];
    }
  } else {
    const double t1 = now();
    // transpose pattern of the lower half -> strictly-upper neighbours, ascending
    ivec tp(n + 1, 0);
    for (int64_t q = 0; q < nnz; ++q) tp[Ai[q] + 1]++;
    for (int j = 0; j < n; ++j) tp[j + 1] += tp[j];
    ivec ti(nnz);
    { ivec f(tp.begin(), tp.end() - 1); for (int j = 0; j < n; ++j) for (int q = Ap[j]; q < Ap[j + 1]; ++q) ti[f[Ai[q]]++] = j; }
    std::vector<int64_t> xadj(n + 1, 0), adj((size_t)std::max<int64_t>(2 * nnz, 1)), pm(n), ipm(n), opts(40);
    for (int i = 0; i < n; ++i) {
      int64_t b = xadj[i];
      for (int q = tp[i]; q < tp[i + 1] - 1; ++q) adj[b++] = ti[q];        // all but the last (= diagonal) entry
      for (int q = Ap[i] + 1; q < Ap[i + 1]; ++q) adj[b++] = Ai[q];        // all but the first (= diagonal) entry
      xadj[i + 1] = b;
    }
    METIS_SetDefaultOptions(opts.data());
    int64_t nn = n;
    const int rc = METIS_NodeND(&nn, xadj.data(), adj.data(), nullptr, opts.data(), pm.data(), ipm.data());
    if (rc != 1) { g_err = "METIS_NodeND failed"; return 5; }
    for (int i = 0; i < n; ++i) perm[i] = (int)pm[i];
    t_ord = now() - t1;
  }

  // ---- analysis of the ordering: etree, postorder, column counts (LSparsity.h:167-247) --------------------
  ivec parent, post, cc;
  {
    Csc U, Lw;
    permute_lower(n, Ap, Ai, nullptr, perm.data(), U, Lw);
    etree_upper(n, U, parent);
    postorder(n, parent, nullptr, post);
    column_counts(n, Lw, parent, post, cc);
  }
  // ---- weighted postorder, combined with the fill-reducing ordering (LSparsity.h:675-715) ------------------
  {
    ivec post2;
    postorder(n, parent, cc.data(), post2);
    ivec np(n), ncc(n), inv(n), npar(n);
    for (int k = 0; k < n; ++k) { np[k] = perm[post2[k]]; ncc[k] = cc[post2[k]]; inv[post2[k]] = k; }
    for (int k = 0; k < n; ++k) { const int op = parent[post2[k]]; npar[k] = op == -1 ? -1 : inv[op]; }
    perm.swap(np); cc.swap(ncc); parent.swap(npar);
  }
  // ---- supernodes and their row patterns (Inspection_BlockC.h:116-879) -------------------------------------
  Csc U, Lw;
  permute_lower(n, Ap, Ai, Ax, perm.data(), U, Lw);
  const int nrelax[3] = {4, 16, 48};
  const double zrelax[3] = {0.8, 0.1, 0.05};
  Supernodes sn;
  find_supernodes(n, parent, cc, nrelax, zrelax, sn);
  const int ns = sn.nsuper;
  std::vector<size_t> pi(ns + 1, 0);
  for (int s = 0; s < ns; ++s) pi[s + 1] = pi[s] + (size_t)sn.snz[s];
  const size_t ssize = pi[ns];
  ivec Ls(std::max<size_t>(ssize, 1));
  {
    std::vector<size_t> fillp(pi.begin(), pi.end() - 1);
    ivec flag(ns, -1);
    int mark = -1;
    for (int s = 0; s < ns; ++s) {
      const int k1 = sn.super[s], k2 = sn.super[s + 1];
      for (int k = k1; k < k2; ++k) Ls[fillp[s]++] = k;
      for (int k = k1; k < k2; ++k) {
        ++mark;
        flag[s] = mark;
        // row k of L: every supernode on the paths from the columns i < k1 of A(:,k) up to s receives row k
        for (int q = U.p[k]; q < U.p[k + 1]; ++q) {
          const int i = U.i[q];
          if (i >= k1) continue;
          for (int si = sn.supermap[i]; flag[si] < mark; si = sn.sparent[si]) {
            if (fillp[si] >= pi[si + 1]) { g_err = "row pattern overflow (inconsistent column counts)"; return 5; }
            Ls[fillp[si]++] = k;
            flag[si] = mark;
          }
        }
      }
    }
    for (int s = 0; s < ns; ++s)
      if (fillp[s] != pi[s + 1]) { g_err = "row pattern does not fill its supernode"; return 5; }
  }
  // ---- executor-facing arrays (LSparsity.h:752-782) ------------------------------------------------------------
  std::vector<size_t> Lp(n + 1, 0), Liptr(n + 1, 0);
  int maxSupWid = 0, maxCol = 0;
  size_t xsize = 0;
  for (int s = 0; s < ns; ++s) {
    const int k1 = sn.super[s], k2 = sn.super[s + 1];
    const size_t len = pi[s + 1] - pi[s];
    maxSupWid = std::max(maxSupWid, k2 - k1);
    maxCol = std::max<int>(maxCol, (int)len);
    for (int j = k1; j < k2; ++j) { Liptr[j] = pi[s]; Lp[j] = xsize + (size_t)(j - k1) * len; }
    xsize += (size_t)(k2 - k1) * len;
  }
  Liptr[n] = pi[ns]; Lp[n] = xsize;
  // ---- node costs and the LBC schedule --------------------------------------------------------------------------
  std::vector<double> nodeCost(ns);
  for (int s = 0; s < ns; ++s)   // computeCostperBlock degenerates to width x rows
    nodeCost[s] = (double)(sn.super[s + 1] - sn.super[s]) * (double)(pi[s + 1] - pi[s]);
  Schedule sch;
  lbc_schedule(ns, sn.sparent, sn.super, nodeCost, costParam, levelParam, divRate, sch);
  if ((int)sch.partition.size() != ns) { g_err = "LBC schedule does not cover every supernode"; return 5; }

  parsy_symbolic* R = new parsy_symbolic();
  memset(R, 0, sizeof(*R));
  R->n = n; R->nsuper = ns; R->nnzA = nnz; R->xsize = (int64_t)xsize; R->ssize = (int64_t)ssize;
  R->maxSupWid = maxSupWid; R->maxCol = maxCol;
  double fl = 0;
  for (int j = 0; j < n; ++j) fl += (double)cc[j] * (double)cc[j];
  R->flops = fl; R->t_ordering = t_ord;
  R->Perm = dup(perm); R->ColCount = dup(cc); R->Parent = dup(parent); R->super = dup(sn.super);
  R->sParent = dup(sn.sparent); R->col2Sup = dup(sn.supermap); R->pi = dup(pi); R->s = dup(Ls); R->p = dup(Lp);
  R->i_ptr = dup(Liptr);
  R->nLevels = (int)sch.levelPtr.size() - 1; R->nParts = (int)sch.parPtr.size() - 1;
  R->levelPtr = dup(sch.levelPtr); R->parPtr = dup(sch.parPtr); R->partition = dup(sch.partition);
  R->A1_p = dup(U.p); R->A1_i = dup(U.i);
  R->A2_p = dup(Lw.p); R->A2_i = dup(Lw.i); R->A2_x = dup(Lw.x); R->A2_src = dup(Lw.src);
  R->t_total = now() - t0;
  *out = R;
  return 0;
}

// Descendants of supernode s exactly as ereach_sn returns them (common/Reach.h:112-143): for every column of s
// and every entry of triu(P A P')(:,k) in stored order, the unmarked path up the supernodal etree is pushed
// in front of what was found before.
extern "C" int parsy_ereach_sn(const parsy_symbolic* sym, int s, int* out) {
  if (!sym || !out || s < 0 || s >= sym->nsuper) return -1;
  const int ns = sym->nsuper;
  static thread_local std::vector<char> marked;
  if ((int)marked.size() != ns) marked.assign(ns, 0);
  ivec stack(ns), path;
  int top = ns;
  const int col1 = sym->super[s], col2 = sym->super[s + 1];
  marked[s] = 1;
  for (int k = col1; k < col2; ++k)
    for (int q = sym->A1_p[k]; q < sym->A1_p[k + 1]; ++q) {
      if (sym->A1_i[q] > k) continue;
      path.clear();
      for (int i = sym->col2Sup[sym->A1_i[q]]; !marked[i]; i = sym->sParent[i]) { path.push_back(i); marked[i] = 1; }
      for (int t = (int)path.size() - 1; t >= 0; --t) stack[--top] = path[t];
    }
  const int cnt = ns - top;
  for (int t = 0; t < cnt; ++t) { out[t] = stack[top + t]; marked[stack[top + t]] = 0; }
  marked[s] = 0;
  return cnt;
}

extern "C" int parsy_etree_level_set(int nsuper, const int* sParent, int* levelPtr, int* levelSet) {
  if (nsuper < 0 || !sParent || !levelPtr || !levelSet) return -1;
  ivec wptr, wset, n2w;
  const int nw = tree_waves(nsuper, sParent, wptr, wset, n2w);
  for (int i = 0; i <= nw; ++i) levelPtr[i] = wptr[i];
  for (int i = 0; i < nsuper; ++i) levelSet[i] = wset[i];
  return nw;
}

extern "C" int64_t parsy_bcsc2csc(const parsy_symbolic* sym, const double* Lx, int* Cp, int* Ci, double* Cx) {
  if (!sym || !Cp) return -1;
  int64_t nz = 0;
  Cp[0] = 0;
  for (int s = 0; s < sym->nsuper; ++s) {
    const int k1 = sym->super[s], k2 = sym->super[s + 1];
    const size_t len = sym->pi[s + 1] - sym->pi[s];
    for (int j = k1; j < k2; ++j) {
      for (size_t t = (size_t)(j - k1); t < len; ++t) {
        if (Ci) Ci[nz] = sym->s[sym->pi[s] + t];
        if (Cx && Lx) Cx[nz] = Lx[sym->p[j] + t];
        ++nz;
      }
      Cp[j + 1] = (int)nz;
    }
  }
  return nz;
}
