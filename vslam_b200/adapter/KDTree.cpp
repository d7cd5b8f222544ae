// KDTree.cpp — implementation of include/KDTree.h over the C ABI. Compiled by the integrator in place of
// the reference's src/KDTree.cpp and linked with libvslam_b200.so.
#include "KDTree.h"

#include <cstdlib>
#include <cstring>
#include <list>
#include <unordered_map>

#include "adapter_common.h"

using namespace vslam_b200_adapter;

void vslam_b200_set_device(int device) {
    Global &g = global();
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.ctx) g.device = device;
}

namespace {

// Device copy of a tree, keyed by the host `root` pointer the caller owns.
struct Entry {
    vb_tree *tree = nullptr;
    uint32_t n = 0;
    std::vector<uint32_t> slot_of;   // original point index -> pre-order slot (value tree only)
    std::list<const void *>::iterator lru;
};

struct Table {
    std::mutex mu;
    std::unordered_map<const void *, Entry> map;
    std::list<const void *> order;   // front = most recently used
    size_t cap = 64;                 // device trees kept alive; older ones are re-imported on demand

    void touch(Entry &e, const void *key) {
        order.erase(e.lru);
        order.push_front(key);
        e.lru = order.begin();
    }
    void drop(const void *key) {
        auto it = map.find(key);
        if (it == map.end()) return;
        vb_kdtree_free(it->second.tree);
        order.erase(it->second.lru);
        map.erase(it);
    }
    Entry &put(const void *key, vb_tree *t, uint32_t n) {
        drop(key);   // a recycled address means the old tree was free()d by its owner
        while (map.size() >= cap) drop(order.back());
        Entry &e = map[key];
        e.tree = t;
        e.n = n;
        order.push_front(key);
        e.lru = order.begin();
        return e;
    }
};

Table &table() {
    static Table t;
    return t;
}

template <typename Node> void link_preorder(Node *nodes, uint32_t n) {
    // slot s of a subtree of length len: left child at s+1 (len/2 nodes), right at s+1+len/2
    struct Item { uint32_t slot, len; };
    std::vector<Item> st;
    if (n) st.push_back({0u, n});
    while (!st.empty()) {
        const Item it = st.back();
        st.pop_back();
        const uint32_t ll = it.len / 2, rl = it.len - ll - 1;
        nodes[it.slot].left = ll ? nodes + it.slot + 1 : NULL;
        nodes[it.slot].right = rl ? nodes + it.slot + 1 + ll : NULL;
        if (ll) st.push_back({it.slot + 1, ll});
        if (rl) st.push_back({it.slot + 1 + ll, rl});
    }
}

std::vector<float> flatten(const std::vector<cv::Point2f> &pts) {
    std::vector<float> f(pts.size() * 2);
    for (size_t i = 0; i < pts.size(); i++) { f[2 * i] = pts[i].x; f[2 * i + 1] = pts[i].y; }
    return f;
}

// Number of nodes hanging off `root`, by walking the links. The struct's own `size` cannot be trusted: the reference (and
// this adapter, to stay observably identical) only ever ADDS to it (src/KDTree.cpp:16), so a struct that was constructed
// twice reports n1 + n2 while `root` holds n2 nodes.
template <typename Node> uint32_t count_nodes(const Node *root) {
    uint32_t n = 0;
    std::vector<const Node *> st;
    if (root) st.push_back(root);
    while (!st.empty()) {
        const Node *nd = st.back();
        st.pop_back();
        n++;
        if (nd->left) st.push_back(nd->left);
        if (nd->right) st.push_back(nd->right);
    }
    return n;
}

// Device tree for a value tree; imports it from the host nodes when the side table does not have it (evicted, or built by
// other code). The nodes of a construct_kdtree result are one pre-order block, so root[0..n) is the whole tree.
Entry &entry_for(const KDTree &t) {
    Table &tb = table();
    auto it = tb.map.find(t.root);
    if (it != tb.map.end()) { tb.touch(it->second, t.root); return it->second; }
    const uint32_t n = count_nodes(t.root);
    std::vector<float> pre((size_t)n * 2);
    for (uint32_t s = 0; s < n; s++) { pre[2 * s] = t.root[s].pt.x; pre[2 * s + 1] = t.root[s].pt.y; }
    vb_tree *dt = nullptr;
    check(vb_kdtree_import(context(), pre.data(), nullptr, n, &dt), "vb_kdtree_import");
    Entry &e = tb.put(t.root, dt, n);   // idx == slot, so slot_of stays empty (identity)
    return e;
}

Entry &entry_for(const frame_kdtree &t, const std::vector<cv::Point2f> &points) {
    Table &tb = table();
    auto it = tb.map.find(t.root);
    if (it != tb.map.end()) { tb.touch(it->second, t.root); return it->second; }
    const uint32_t n = count_nodes(t.root);
    std::vector<float> pre((size_t)n * 2);
    std::vector<uint32_t> idx(n);
    for (uint32_t s = 0; s < n; s++) {
        const usize pi = t.root[s].pt_index;
        idx[s] = (uint32_t)pi;
        pre[2 * s] = points[pi].x;
        pre[2 * s + 1] = points[pi].y;
    }
    vb_tree *dt = nullptr;
    check(vb_kdtree_import(context(), pre.data(), idx.data(), n, &dt), "vb_kdtree_import");
    return tb.put(t.root, dt, n);
}

// CSR radius query with retry on capacity
void radius_csr(vb_tree *dt, const std::vector<float> &q, uint32_t nq, float radius, std::vector<uint32_t> &off,
                std::vector<uint32_t> &out) {
    off.assign(nq + 1, 0);
    uint64_t cap = std::max<uint64_t>(64, (uint64_t)nq * 16), total = 0;
    for (;;) {
        out.resize(cap);
        const int rc = vb_kdtree_radius(dt, q.data(), nq, radius, off.data(), out.data(), cap, &total);
        if (rc == VB_ERR_CAPACITY && total > cap) { cap = total; continue; }
        check(rc, "vb_kdtree_radius");
        out.resize(total);
        return;
    }
}

}  // namespace

void vslam_b200_kdtree_release(const void *root) {
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    tb.drop(root);
}

// ---------------------------------------------------------------------------------------------------
// value tree (reference src/KDTree.cpp:25-35)
void construct_kdtree(KDTree &kdtree, const std::vector<cv::Point2f> &points) {
    const uint32_t n = (uint32_t)points.size();
    if (n == 0) {
        kdtree.root = NULL;
        return;
    }
    vb_tree *dt = nullptr;
    const std::vector<float> flat = flatten(points);
    check(vb_kdtree_build(context(), flat.data(), n, &dt), "vb_kdtree_build");
    std::vector<uint32_t> idx(n);
    std::vector<float> pre((size_t)n * 2);
    check(vb_kdtree_export(dt, idx.data(), pre.data()), "vb_kdtree_export");
    kdtree.root = (KDTree::KDTreeNode *)malloc((size_t)n * sizeof(KDTree::KDTreeNode));   // caller free()s it
    for (uint32_t s = 0; s < n; s++) kdtree.root[s].pt = cv::Point2f(pre[2 * s], pre[2 * s + 1]);
    link_preorder(kdtree.root, n);
    kdtree.size += n;   // the reference advances size once per node and never resets it (:16)
    kdtree.height = (u8)(std::floor(std::log2((double)n)) + 1);
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    Entry &e = tb.put(kdtree.root, dt, n);
    e.slot_of.assign(n, 0);
    for (uint32_t s = 0; s < n; s++) e.slot_of[idx[s]] = s;
}

std::vector<cv::Point2f> nearest_batch(const KDTree &kdtree, const std::vector<cv::Point2f> &queries, float max_distance_sq) {
    std::vector<cv::Point2f> res(queries.size());   // default {0,0}: what the reference returns on a miss
    if (queries.empty() || kdtree.root == NULL) return res;
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    Entry &e = entry_for(kdtree);
    const std::vector<float> q = flatten(queries);
    std::vector<float> out(q.size());
    check(vb_kdtree_nearest(e.tree, q.data(), (uint32_t)queries.size(), max_distance_sq, out.data(), nullptr, nullptr),
          "vb_kdtree_nearest");
    for (size_t i = 0; i < queries.size(); i++) res[i] = cv::Point2f(out[2 * i], out[2 * i + 1]);
    return res;
}

cv::Point2f nearest(const KDTree &kdtree, const cv::Point2f &query_pt, float max_distance_sq) {
    return nearest_batch(kdtree, std::vector<cv::Point2f>(1, query_pt), max_distance_sq)[0];
}

std::vector<std::vector<cv::Point2f> > radius_search_batch(const KDTree &kdtree, const std::vector<cv::Point2f> &queries,
                                                           float radius) {
    std::vector<std::vector<cv::Point2f> > res(queries.size());
    if (queries.empty() || kdtree.root == NULL) return res;
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    Entry &e = entry_for(kdtree);
    std::vector<uint32_t> off, out;
    radius_csr(e.tree, flatten(queries), (uint32_t)queries.size(), radius, off, out);
    for (size_t i = 0; i < queries.size(); i++) {
        res[i].reserve(off[i + 1] - off[i]);
        for (uint32_t k = off[i]; k < off[i + 1]; k++) {
            const uint32_t slot = e.slot_of.empty() ? out[k] : e.slot_of[out[k]];
            res[i].push_back(kdtree.root[slot].pt);
        }
    }
    return res;
}

std::vector<cv::Point2f> radius_search(const KDTree &kdtree, const cv::Point2f &query_pt, float radius) {
    return radius_search_batch(kdtree, std::vector<cv::Point2f>(1, query_pt), radius)[0];
}

// ---------------------------------------------------------------------------------------------------
// index tree (reference src/KDTree.cpp:107-121, :145-150)
void construct_kdtree(frame_kdtree &kdtree, const std::vector<cv::Point2f> &points) {
    const uint32_t n = (uint32_t)points.size();
    if (n == 0) {
        kdtree.root = NULL;
        return;
    }
    vb_tree *dt = nullptr;
    const std::vector<float> flat = flatten(points);
    check(vb_kdtree_build(context(), flat.data(), n, &dt), "vb_kdtree_build");
    std::vector<uint32_t> idx(n);
    check(vb_kdtree_export(dt, idx.data(), nullptr), "vb_kdtree_export");
    kdtree.root = (frame_kdtree::KDTreeNode *)malloc((size_t)n * sizeof(frame_kdtree::KDTreeNode));
    for (uint32_t s = 0; s < n; s++) kdtree.root[s].pt_index = idx[s];
    link_preorder(kdtree.root, n);
    kdtree.size += n;
    kdtree.height = (u8)(std::floor(std::log2((double)n)) + 1);
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    tb.put(kdtree.root, dt, n);
}

std::vector<std::vector<usize> > radius_search_batch(const frame_kdtree &kdtree, const std::vector<cv::Point2f> &points,
                                                     const std::vector<cv::Point2f> &queries, float radius) {
    std::vector<std::vector<usize> > res(queries.size());
    if (queries.empty() || kdtree.root == NULL) return res;
    Table &tb = table();
    std::lock_guard<std::mutex> lk(tb.mu);
    Entry &e = entry_for(kdtree, points);
    std::vector<uint32_t> off, out;
    radius_csr(e.tree, flatten(queries), (uint32_t)queries.size(), radius, off, out);
    for (size_t i = 0; i < queries.size(); i++) res[i].assign(out.begin() + off[i], out.begin() + off[i + 1]);
    return res;
}

std::vector<usize> radius_search(const frame_kdtree kdtree, const std::vector<cv::Point2f> &points,
                                 const cv::Point2f &query_pt, float radius) {
    return radius_search_batch(kdtree, points, std::vector<cv::Point2f>(1, query_pt), radius)[0];
}
