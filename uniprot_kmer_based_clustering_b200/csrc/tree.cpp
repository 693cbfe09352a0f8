// tree.cpp — host-side incremental tree clustering (kc_tree_*, include/kc_host.h).
//
// Restates the live code of the reference's src/tree.rs (which the reference never compiled:
// `// mod tree;` at src/main.rs:15) over the engine's per-protein id lists:
//   Node::new_leaf src/tree.rs:64-106, clone_and_clean :151-177, balance :179-265,
//   add_child :267-385, Tree::new / add_protein :519-536.
// The two Protein accessors tree.rs calls but protein.rs lacks are taken as in SURVEY.md §8c:
// get_five_hash() = the protein's repeated-k-mer ids, get_five_hash_map() = their bit-array.
// The north star keeps this stage on the host.  What changes against a literal port:
//   * c sets (complete intersections) are sorted id vectors, u sets (unions) are bitsets for
//     internal nodes, so add_child costs O(|ids|) instead of O(vocabulary);
//   * balance() re-computes |c_i ∩ c_j| for EVERY pair of children on every call in the
//     reference; here the pairwise similarities of a node's children are cached and only the
//     rows of new or changed children are recomputed (same values, same loop-order tie-breaking).
// Specification: a literal Python model kept with the test infrastructure (tests/test_tree_host.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/kc_b200.h"
#include "../../include/kc_host.h"

namespace {

struct TNode {
  std::vector<int> children;
  std::vector<uint32_t> c;        // complete intersection of the subtree's id sets (sorted)
  std::vector<uint32_t> u_ids;    // union as a sorted list (leaves only)
  std::vector<uint64_t> u_bits;   // union as a bitset (internal nodes)
  int protein = -1;
  // cached similarities of the children: sim[i][j] = |c_i ∩ c_j| for j < i
  std::vector<std::vector<uint32_t>> sim;
  std::vector<char> dirty;        // child i needs its similarities recomputed
};

uint32_t intersect_count(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  size_t i = 0, j = 0;
  uint32_t n = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i] < b[j]) ++i;
    else if (a[i] > b[j]) ++j;
    else {
      ++n;
      ++i;
      ++j;
    }
  }
  return n;
}

}  // namespace

struct kc_tree {
  std::vector<TNode> nodes;
  int root = -1;
  uint32_t n_ids = 0;
  uint64_t n_proteins = 0, n_merges = 0, n_no_common = 0;
  size_t words() const { return (n_ids + 63) / 64; }

  int new_leaf(int protein, const uint32_t* ids, uint64_t n) {
    nodes.emplace_back();
    TNode& t = nodes.back();
    t.protein = protein;
    t.c.assign(ids, ids + n);
    t.u_ids = t.c;
    return (int)nodes.size() - 1;
  }

  bool has_common(const TNode& a, const TNode& b) const {  // a.u ∩ b.u non-empty
    const bool ab = !a.u_bits.empty(), bb = !b.u_bits.empty();
    if (ab && bb) {
      for (size_t w = 0; w < a.u_bits.size(); ++w)
        if (a.u_bits[w] & b.u_bits[w]) return true;
      return false;
    }
    if (ab || bb) {
      const TNode& bits = ab ? a : b;
      const TNode& list = ab ? b : a;
      for (uint32_t id : list.u_ids)
        if (bits.u_bits[id >> 6] >> (id & 63) & 1) return true;
      return false;
    }
    return intersect_count(a.u_ids, b.u_ids) > 0;
  }

  void unite_into(TNode& dst, const TNode& src) {  // dst.u |= src.u (dst becomes or stays internal)
    if (dst.u_bits.empty()) {
      dst.u_bits.assign(words(), 0);
      for (uint32_t id : dst.u_ids) dst.u_bits[id >> 6] |= 1ull << (id & 63);
      dst.u_ids.clear();
      dst.u_ids.shrink_to_fit();
    }
    if (!src.u_bits.empty())
      for (size_t w = 0; w < dst.u_bits.size(); ++w) dst.u_bits[w] |= src.u_bits[w];
    else
      for (uint32_t id : src.u_ids) dst.u_bits[id >> 6] |= 1ull << (id & 63);
  }

  static void intersect_into(std::vector<uint32_t>& dst, const std::vector<uint32_t>& src) {
    size_t i = 0, j = 0, w = 0;
    while (i < dst.size() && j < src.size()) {
      if (dst[i] < src[j]) ++i;
      else if (dst[i] > src[j]) ++j;
      else {
        dst[w++] = dst[i];
        ++i;
        ++j;
      }
    }
    dst.resize(w);
  }

  void push_child(int parent, int child) {
    TNode& p = nodes[parent];
    p.children.push_back(child);
    p.sim.emplace_back(p.children.size() - 1, 0u);
    p.dirty.push_back(1);
  }

  void remove_child(int parent, size_t idx) {
    TNode& p = nodes[parent];
    p.children.erase(p.children.begin() + idx);
    p.sim.erase(p.sim.begin() + idx);
    for (size_t i = idx; i < p.sim.size(); ++i) p.sim[i].erase(p.sim[i].begin() + idx);
    p.dirty.erase(p.dirty.begin() + idx);
  }

  void refresh_sims(int parent) {
    TNode& p = nodes[parent];
    const size_t m = p.children.size();
    for (size_t i = 0; i < m; ++i) {
      if (!p.dirty[i]) continue;
      for (size_t j = 0; j < m; ++j) {
        if (j == i) continue;
        if (j > i && p.dirty[j]) continue;  // the pair is done when j's turn comes
        const uint32_t s = intersect_count(nodes[p.children[i]].c, nodes[p.children[j]].c);
        if (j < i) p.sim[i][j] = s; else p.sim[j][i] = s;
      }
    }
    std::fill(p.dirty.begin(), p.dirty.end(), 0);
  }

  // Node::balance, src/tree.rs:179-265
  void balance(int curr) {
    refresh_sims(curr);
    const TNode& p = nodes[curr];
    uint32_t best = 0;
    size_t bi = 0, bj = 0;
    bool have_min = false;
    uint32_t mn = 0;
    for (size_t i = 1; i < p.children.size(); ++i)
      for (size_t j = 0; j < i; ++j) {
        const uint32_t s = p.sim[i][j];
        if (s > best) {
          best = s;
          bi = i;
          bj = j;
        }
        if (!have_min || mn > s) {
          mn = s;
          have_min = true;
        }
      }
    if (!have_min || best <= mn) return;
    ++n_merges;
    const int one = p.children[bi], two = p.children[bj];
    if (nodes[one].children.size() < nodes[two].children.size()) {
      remove_child(curr, bj);
      add_child(one, two);
      mark_dirty(curr, one);
    } else {
      remove_child(curr, bi);
      add_child(two, one);
      mark_dirty(curr, two);
    }
  }

  void mark_dirty(int parent, int child) {  // the child's c changed: its cached similarities are stale
    TNode& p = nodes[parent];
    for (size_t i = 0; i < p.children.size(); ++i)
      if (p.children[i] == child) p.dirty[i] = 1;
  }

  // Node::add_child, src/tree.rs:267-385
  void add_child(int curr, int child) {
    if (nodes[curr].children.empty()) {
      // clone_and_clean: the leaf's content moves into a fresh node, curr becomes internal
      nodes.emplace_back();
      const int cloned = (int)nodes.size() - 1;
      {
        TNode& cl = nodes[cloned];
        TNode& cu = nodes[curr];
        cl.c = cu.c;
        cl.u_ids = cu.u_ids;
        cl.u_bits = cu.u_bits;
        cl.protein = cu.protein;
        cu.protein = -1;
      }
      unite_into(nodes[curr], nodes[child]);
      intersect_into(nodes[curr].c, nodes[child].c);
      push_child(curr, cloned);
      if (nodes[child].children.empty()) {
        push_child(curr, child);
      } else {
        const std::vector<int> grand = nodes[child].children;  // the child node itself is dropped
        for (int g : grand) push_child(curr, g);
      }
    } else {
      const bool common = has_common(nodes[curr], nodes[child]);
      unite_into(nodes[curr], nodes[child]);
      intersect_into(nodes[curr].c, nodes[child].c);
      push_child(curr, child);
      if (common) balance(curr);
      else ++n_no_common;
    }
  }
};

extern "C" {

int kc_tree_build(const uint64_t* row_offsets, const uint32_t* ids, uint64_t n_proteins, uint32_t n_ids,
                  kc_tree** out) {
  if (!out || !row_offsets || (n_proteins && row_offsets[n_proteins] && !ids)) return KC_EINVAL;
  kc_tree* t = new kc_tree();
  t->n_ids = n_ids;
  t->n_proteins = n_proteins;
  t->nodes.reserve(2 * n_proteins + 4);
  for (uint64_t p = 0; p < n_proteins; ++p) {
    const uint64_t lo = row_offsets[p], hi = row_offsets[p + 1];
    for (uint64_t i = lo; i < hi; ++i)
      if (ids[i] >= n_ids || (i > lo && ids[i] <= ids[i - 1])) {
        delete t;
        return KC_EINVAL;  // rows must be ascending ids below n_ids
      }
    const int leaf = t->new_leaf((int)p, ids + lo, hi - lo);
    if (p == 0) t->root = leaf;                 // Tree::new
    else t->add_child(t->root, leaf);           // Tree::add_protein
  }
  *out = t;
  return KC_OK;
}

void kc_tree_free(kc_tree* t) { delete t; }
uint64_t kc_tree_n_merges(const kc_tree* t) { return t->n_merges; }
uint64_t kc_tree_n_no_common(const kc_tree* t) { return t->n_no_common; }

// Preorder serialisation: a leaf is its protein index (>= 0); an internal node is -(number of
// children) followed by its children in order.  Returns the token count (call with out = NULL
// to size the buffer).
uint64_t kc_tree_serialize(const kc_tree* t, int64_t* out, uint64_t capacity) {
  if (t->root < 0) return 0;
  uint64_t n = 0;
  std::vector<int> stack{t->root};
  while (!stack.empty()) {
    const int v = stack.back();
    stack.pop_back();
    const TNode& nd = t->nodes[v];
    const int64_t tok = nd.children.empty() ? (int64_t)nd.protein : -(int64_t)nd.children.size();
    if (out && n < capacity) out[n] = tok;
    ++n;
    for (size_t i = nd.children.size(); i-- > 0;) stack.push_back(nd.children[i]);
  }
  return n;
}

// cluster_of[p] = index of the root child (top-level cluster) that holds protein p
int kc_tree_clusters(const kc_tree* t, uint32_t* cluster_of, uint32_t* n_clusters) {
  if (!t || !n_clusters || (t->n_proteins && !cluster_of)) return KC_EINVAL;
  if (t->root < 0) {
    *n_clusters = 0;
    return KC_OK;
  }
  const TNode& root = t->nodes[t->root];
  if (root.children.empty()) {
    cluster_of[root.protein] = 0;
    *n_clusters = 1;
    return KC_OK;
  }
  for (size_t ci = 0; ci < root.children.size(); ++ci) {
    std::vector<int> stack{root.children[ci]};
    while (!stack.empty()) {
      const int v = stack.back();
      stack.pop_back();
      const TNode& nd = t->nodes[v];
      if (nd.children.empty()) cluster_of[nd.protein] = (uint32_t)ci;
      for (int ch : nd.children) stack.push_back(ch);
    }
  }
  *n_clusters = (uint32_t)root.children.size();
  return KC_OK;
}

}  // extern "C"
