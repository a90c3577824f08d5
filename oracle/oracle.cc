/*
 * oracle.cc -- single-threaded CPU restatement of the scan-insertion hot path of
 * ethz-asl/volumetric_mapping.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  It is the parity checker for the CUDA path and the
 * timed CPU baseline; the product never links, loads or calls it.
 *
 * PARITY UNPINNED by the reference: it ships no tests / golden vectors, and none of octomap,
 * PCL, OpenCV (C++), Eigen or minkindr exist in this environment, so the reference cannot be
 * compiled.  What pins this file: cv2 4.13 `reprojectImageTo3D` golden vectors
 * (tests/golden/reproject_*.npz, generator tests/golden/make_golden.py) and the hand-derived
 * known-answer tests in tests/test_oracle_*.py.
 *
 * In-tree code followed (all paths relative to /root/reference):
 *   octomap_world/src/octomap_world.cc:84-104   setOctomapParameters  -> World::set_params
 *   octomap_world/src/octomap_world.cc:110-141  insertPointcloudIntoMapImpl
 *   octomap_world/src/octomap_world.cc:143-175  insertProjectedDisparityIntoMapImpl
 *   octomap_world/src/octomap_world.cc:177-230  castRay
 *   octomap_world/src/octomap_world.cc:232-236  isValidPoint
 *   octomap_world/src/octomap_world.cc:238-264  updateOccupancy
 *   octomap_world/src/octomap_world.cc:346-360  getCellStatusPoint
 *   octomap_world/src/octomap_world.cc:395-418  getLineStatus
 *   volumetric_map_base/src/world_base.cc:51-87 insertDisparityImage (Q fix-up + reprojection)
 *
 * Un-vendored third-party behaviour restated (not present under /root/reference; versions
 * are not pinned by the reference's package.xml files -- assumed versions below):
 *   octomap 1.9-series : OcTreeKey / KeyHash / KeyRay, coordToKey(Checked), keyToCoord,
 *                        computeRayKeys, search, updateNode (+ early abort, expand, prune,
 *                        inner max), updateInnerOccupancy, logodds(), math::Vector3 float ops.
 *   PCL 1.7 (scalar)   : removeNaNFromPointCloud, transformPointCloud(Matrix4d), dense branch.
 *   OpenCV 4.13        : reprojectImageTo3D(handleMissingValues=true).
 *   Eigen 3 / minkindr : Quaternion::toRotationMatrix, Quaternion::_transformVector,
 *                        QuatTransformation::operator*(Vector3).
 *
 * Build: g++ -O3 -ffp-contract=off (no -march=native, no -ffast-math) -- see oracle/Makefile.
 * The data structures deliberately keep the reference's shape (std::unordered_set key sets
 * with octomap's hash, pointer octree with per-node child arrays, non-lazy updateNode with
 * prune + inner max on unwind, full updateInnerOccupancy per scan, 100000-entry KeyRay
 * constructed per getLineStatus call): this file is also the CPU baseline that is timed.
 */
#include "oracle.h"

#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <unordered_set>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// octomap::OcTreeKey, KeyHash, KeySet, KeyRay  [octomap OcTreeKey.h]
// ---------------------------------------------------------------------------------------
typedef uint16_t key_type;

struct Key {
  key_type k[3];
  Key() { k[0] = k[1] = k[2] = 0; }
  Key(key_type a, key_type b, key_type c) { k[0] = a; k[1] = b; k[2] = c; }
  bool operator==(const Key& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
  bool operator!=(const Key& o) const { return !(*this == o); }
};

struct KeyHash {
  size_t operator()(const Key& key) const {
    // octomap 1.9 OcTreeKey::KeyHash.  Only affects iteration order, never results.
    return static_cast<size_t>(key.k[0]) + 1447 * static_cast<size_t>(key.k[1]) +
           345637 * static_cast<size_t>(key.k[2]);
  }
};
typedef std::unordered_set<Key, KeyHash> KeySet;

// octomap::KeyRay: fixed 100000-entry vector allocated on construction, cursor reset().
struct KeyRay {
  static const size_t kMaxSize = 100000;
  std::vector<Key> ray;
  size_t end_of_ray;
  KeyRay() : ray(kMaxSize), end_of_ray(0) {}
  void reset() { end_of_ray = 0; }
  void addKey(const Key& k) { ray[end_of_ray++] = k; }
  size_t size() const { return end_of_ray; }
};

// ---------------------------------------------------------------------------------------
// octomath::Vector3 (3 floats; all arithmetic in float)  [octomap math/Vector3.h]
// ---------------------------------------------------------------------------------------
struct Vec3f {
  float d[3];
  Vec3f() { d[0] = d[1] = d[2] = 0.f; }
  Vec3f(float x, float y, float z) { d[0] = x; d[1] = y; d[2] = z; }
  float operator()(int i) const { return d[i]; }
  float& operator()(int i) { return d[i]; }
  Vec3f operator-(const Vec3f& o) const { return Vec3f(d[0] - o.d[0], d[1] - o.d[1], d[2] - o.d[2]); }
  Vec3f operator+(const Vec3f& o) const { return Vec3f(d[0] + o.d[0], d[1] + o.d[1], d[2] + o.d[2]); }
  Vec3f operator*(float x) const { return Vec3f(d[0] * x, d[1] * x, d[2] * x); }
  void operator/=(float x) { d[0] /= x; d[1] /= x; d[2] /= x; }
  // norm_sq(): products and sums in float, left to right; result widened to double.
  double norm_sq() const { return (double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); }
  double norm() const { return std::sqrt(norm_sq()); }
  Vec3f normalized() const {
    Vec3f r(*this);
    double len = r.norm();
    if (len > 0) r /= (float)len;
    return r;
  }
};

// x86-64 `(int)floor(v)` compiles to cvttsd2si, which yields INT_MIN ("integer indefinite")
// for NaN and for anything outside int range.  The C++ cast is UB there; the reference was
// only ever run on x86, so that is the behaviour pinned here (and mirrored on the GPU).
inline int x86_double_to_int(double v) {
  if (!(v >= -2147483648.0 && v < 2147483648.0)) return INT_MIN;
  return (int)v;
}

// ---------------------------------------------------------------------------------------
// Occupancy octree  [octomap OcTreeBaseImpl.hxx, OccupancyOcTreeBase.hxx, OcTreeNode.h]
// ---------------------------------------------------------------------------------------
struct Node {
  Node** children;  // NULL, or array of 8 (entries may be NULL) -- octomap 1.8+ layout
  float value;      // log-odds
  Node() : children(NULL), value(0.0f) {}
};

class OcTree {
 public:
  static const unsigned tree_depth = 16;
  static const unsigned tree_max_val = 32768;

  explicit OcTree(double res)
      : root(NULL), tree_size(0), resolution(res), resolution_factor(1.0 / res) {
    // OccupancyOcTreeBase defaults (overwritten by setOctomapParameters right away).
    setProbHit(0.7);
    setProbMiss(0.4);
    setClampingThresMin(0.1192);
    setClampingThresMax(0.971);
    setOccupancyThres(0.5);
  }
  ~OcTree() { clear(); }

  static float logodds(double p) { return (float)std::log(p / (1 - p)); }
  void setProbHit(double p) { prob_hit_log = logodds(p); }
  void setProbMiss(double p) { prob_miss_log = logodds(p); }
  void setClampingThresMin(double p) { clamping_thres_min = logodds(p); }
  void setClampingThresMax(double p) { clamping_thres_max = logodds(p); }
  void setOccupancyThres(double p) { occ_prob_thres_log = logodds(p); }
  double getResolution() const { return resolution; }

  void clear() {
    if (root) {
      deleteNodeRecurs(root);
      root = NULL;
      tree_size = 0;
    }
  }
  size_t size() const { return tree_size; }

  // --- key <-> coordinate -------------------------------------------------------------
  key_type coordToKey(double coordinate) const {
    return (key_type)(x86_double_to_int(std::floor(resolution_factor * coordinate)) +
                      (int)tree_max_val);
  }
  Key coordToKey(const Vec3f& p) const {
    return Key(coordToKey((double)p(0)), coordToKey((double)p(1)), coordToKey((double)p(2)));
  }
  bool coordToKeyChecked(double coordinate, key_type& keyval) const {
    int scaled_coord =
        x86_double_to_int(std::floor(resolution_factor * coordinate)) + (int)tree_max_val;
    if ((scaled_coord >= 0) && (((unsigned int)scaled_coord) < (2 * tree_max_val))) {
      keyval = (key_type)scaled_coord;
      return true;
    }
    return false;
  }
  bool coordToKeyChecked(double x, double y, double z, Key& key) const {
    if (!(coordToKeyChecked(x, key.k[0]) && coordToKeyChecked(y, key.k[1]) &&
          coordToKeyChecked(z, key.k[2])))
      return false;
    return true;
  }
  bool coordToKeyChecked(const Vec3f& p, Key& key) const {
    for (unsigned i = 0; i < 3; i++)
      if (!coordToKeyChecked((double)p(i), key.k[i])) return false;
    return true;
  }
  double keyToCoord(key_type key) const {
    return (double((int)key - (int)tree_max_val) + 0.5) * resolution;
  }
  Vec3f keyToCoord(const Key& key) const {
    return Vec3f(float(keyToCoord(key.k[0])), float(keyToCoord(key.k[1])),
                 float(keyToCoord(key.k[2])));
  }

  // --- computeRayKeys (Amanatides & Woo 3-D DDA in key space) -----------------------------
  bool computeRayKeys(const Vec3f& origin, const Vec3f& end, KeyRay& ray) const {
    ray.reset();
    Key key_origin, key_end;
    if (!coordToKeyChecked(origin, key_origin) || !coordToKeyChecked(end, key_end)) return false;
    if (key_origin == key_end) return true;  // same cell: empty ray
    ray.addKey(key_origin);

    Vec3f direction = (end - origin);
    float length = (float)direction.norm();
    direction /= length;

    int step[3];
    double tMax[3];
    double tDelta[3];
    Key current_key = key_origin;

    for (unsigned i = 0; i < 3; ++i) {
      if (direction(i) > 0.0) step[i] = 1;
      else if (direction(i) < 0.0) step[i] = -1;
      else step[i] = 0;

      if (step[i] != 0) {
        double voxelBorder = keyToCoord(current_key.k[i]);
        voxelBorder += (float)(step[i] * resolution * 0.5);
        tMax[i] = (voxelBorder - (double)origin(i)) / (double)direction(i);
        tDelta[i] = resolution / std::fabs((double)direction(i));
      } else {
        tMax[i] = std::numeric_limits<double>::max();
        tDelta[i] = std::numeric_limits<double>::max();
      }
    }

    bool done = false;
    while (!done) {
      unsigned dim;
      if (tMax[0] < tMax[1]) {
        if (tMax[0] < tMax[2]) dim = 0; else dim = 2;
      } else {
        if (tMax[1] < tMax[2]) dim = 1; else dim = 2;
      }
      current_key.k[dim] = (key_type)(current_key.k[dim] + step[dim]);
      tMax[dim] += tDelta[dim];

      if (current_key == key_end) {
        done = true;
        break;
      } else {
        double dist_from_origin = std::min(std::min(tMax[0], tMax[1]), tMax[2]);
        if (dist_from_origin > length) {
          done = true;
          break;
        } else {
          // octomap asserts ray.size() < sizeMax-1 (UB in release builds); the oracle and
          // the GPU walker both stop a ray that would overflow the KeyRay instead.
          if (ray.size() >= KeyRay::kMaxSize - 1) { done = true; break; }
          ray.addKey(current_key);
        }
      }
    }
    return true;
  }

  // --- search --------------------------------------------------------------------------
  static unsigned computeChildIdx(const Key& key, int depth) {
    unsigned pos = 0;
    if (key.k[0] & (1 << depth)) pos += 1;
    if (key.k[1] & (1 << depth)) pos += 2;
    if (key.k[2] & (1 << depth)) pos += 4;
    return pos;
  }
  static bool nodeChildExists(const Node* n, unsigned i) {
    return n->children != NULL && n->children[i] != NULL;
  }
  static bool nodeHasChildren(const Node* n) {
    if (n->children == NULL) return false;
    for (unsigned i = 0; i < 8; i++)
      if (n->children[i] != NULL) return true;
    return false;
  }
  Node* search(const Key& key) const {
    if (root == NULL) return NULL;
    Node* curNode = root;
    for (int i = (int)tree_depth - 1; i >= 0; --i) {
      unsigned pos = computeChildIdx(key, i);
      if (nodeChildExists(curNode, pos)) {
        curNode = curNode->children[pos];
      } else {
        if (!nodeHasChildren(curNode)) return curNode;  // pruned leaf
        return NULL;
      }
    }
    return curNode;
  }
  Node* search(double x, double y, double z) const {
    Key key;
    if (!coordToKeyChecked(x, y, z, key)) return NULL;  // (octomap prints an error)
    return search(key);
  }
  bool isNodeOccupied(const Node* n) const { return n->value >= occ_prob_thres_log; }

  // --- node geometry at a depth  [octomap OcTreeBaseImpl.h keyToCoord(key, depth), getNodeSize] ----
  double getNodeSize(unsigned depth) const { return resolution * double(1 << (tree_depth - depth)); }
  double keyToCoord(key_type key, unsigned depth) const {
    if (depth == 0) return 0.0;
    if (depth == tree_depth) return keyToCoord(key);
    return (std::floor((double(key) - double(tree_max_val)) / double(1 << (tree_depth - depth))) + 0.5) *
           getNodeSize(depth);
  }
  Node* search(const Vec3f& p) const {           // search(const point3d&): checked key, NULL outside
    Key key;
    if (!coordToKeyChecked(p, key)) return NULL;
    return search(key);
  }

  // --- leaf_bbx_iterator  [octomap OcTreeIterator.hxx] ---------------------------------------
  // Visits, depth first in child order, every leaf (a node without children, or a node at the
  // finest depth) whose key cube passes the iterator's overlap test against [minKey, maxKey].
  // The test is made when a child is pushed and is conservative by one key on the upper side:
  //   minKey <= key + (tree_max_val >> depth)  &&  maxKey >= key - (tree_max_val >> depth).
  // `f(node, key, depth)` returns false to stop the iteration; the function then returns false.
  static void computeChildKey(unsigned pos, key_type center_offset_key, const Key& parent_key, Key& child_key) {
    for (unsigned a = 0; a < 3; a++) {
      if (pos & (1u << a)) child_key.k[a] = (key_type)(parent_key.k[a] + center_offset_key);
      else child_key.k[a] = (key_type)(parent_key.k[a] - center_offset_key - (center_offset_key ? 0 : 1));
    }
  }
  template <class F>
  bool leafsBBXRecurs(const Node* node, const Key& key, unsigned depth, const Key& minKey, const Key& maxKey,
                      F& f) const {
    if (depth == tree_depth || !nodeHasChildren(node)) return f(node, key, depth);
    const key_type center_offset_key = (key_type)(tree_max_val >> (depth + 1));
    for (unsigned i = 0; i < 8; i++) {
      if (!nodeChildExists(node, i)) continue;
      Key ck;
      computeChildKey(i, center_offset_key, key, ck);
      bool overlap = true;
      for (unsigned a = 0; a < 3; a++)
        overlap = overlap && ((int)minKey.k[a] <= (int)ck.k[a] + (int)center_offset_key) &&
                  ((int)maxKey.k[a] >= (int)ck.k[a] - (int)center_offset_key);
      if (overlap && !leafsBBXRecurs(node->children[i], ck, depth + 1, minKey, maxKey, f)) return false;
    }
    return true;
  }
  template <class F>
  bool forEachLeafBBX(const Vec3f& min, const Vec3f& max, F f) const {
    if (root == NULL) return true;
    Key minKey, maxKey;
    if (!coordToKeyChecked(min, minKey) || !coordToKeyChecked(max, maxKey)) return true;   // end iterator
    return leafsBBXRecurs(root, Key(tree_max_val, tree_max_val, tree_max_val), 0, minKey, maxKey, f);
  }

  // --- getUnknownLeafCenters(pmin, pmax, depth = 0)  [octomap OcTreeBaseImpl.hxx] ----------------
  // Only "is there any" is needed by the caller.  The upstream loop adds the step BEFORE each
  // query, in float: the layer of pmin's voxel is never tested, the layer `steps` voxels above is.
  bool anyUnknownLeafCenter(const Vec3f& pmin, const Vec3f& pmax) const {
    const Vec3f pmin_clamped = keyToCoord(coordToKey(pmin));
    const Vec3f pmax_clamped = keyToCoord(coordToKey(pmax));
    float diff[3];
    unsigned steps[3];
    const float step_size = (float)(resolution * std::pow(2.0, 0.0));
    for (int i = 0; i < 3; ++i) {
      diff[i] = pmax_clamped(i) - pmin_clamped(i);
      const float q = std::floor(diff[i] / step_size);
      steps[i] = q > 0.0f ? (unsigned)q : 0u;   // (negative: the box is inverted, nothing to scan)
    }
    Vec3f p = pmin_clamped;
    for (unsigned x = 0; x < steps[0]; ++x) {
      p.d[0] += step_size;
      for (unsigned y = 0; y < steps[1]; ++y) {
        p.d[1] += step_size;
        for (unsigned z = 0; z < steps[2]; ++z) {
          p.d[2] += step_size;
          if (search(p) == NULL) return true;
        }
        p.d[2] = pmin_clamped(2);
      }
      p.d[1] = pmin_clamped(1);
    }
    return false;
  }

  // --- updateNode ----------------------------------------------------------------------
  Node* updateNode(const Key& key, bool occupied) {
    float logOdds = prob_miss_log;
    if (occupied) logOdds = prob_hit_log;
    return updateNode(key, logOdds);
  }
  Node* updateNode(const Key& key, float log_odds_update) {
    // early abort (no change will happen)
    Node* leaf = search(key);
    if (leaf && ((log_odds_update >= 0 && leaf->value >= clamping_thres_max) ||
                 (log_odds_update <= 0 && leaf->value <= clamping_thres_min))) {
      return leaf;
    }
    bool createdRoot = false;
    if (root == NULL) {
      root = new Node();
      tree_size++;
      createdRoot = true;
    }
    return updateNodeRecurs(root, createdRoot, key, 0, log_odds_update);
  }
  // setNodeValue is not on the path; leaves are only ever written through updateNode.

  void updateInnerOccupancy() {
    if (root) updateInnerOccupancyRecurs(root, 0);
  }
  // OccupancyOcTreeBase::setNodeValue(point3d, float log_odds, lazy_eval = true): out-of-bounds
  // points are ignored; the value is clamped to [clamping_thres_min, clamping_thres_max] ("clamp
  // log odds within range" at the top of setNodeValue(key, ...)); inner nodes are left stale
  // (updateInnerOccupancy follows)
  void setNodeValueLazy(const Vec3f& p, float log_odds_value) {
    Key key;
    if (!coordToKeyChecked(p, key)) return;
    log_odds_value = std::min(std::max(log_odds_value, clamping_thres_min), clamping_thres_max);
    bool just_created = false;
    if (root == NULL) {
      root = new Node();
      tree_size++;
      just_created = true;
    }
    Node* node = root;
    for (unsigned depth = 0; depth < tree_depth; depth++) {
      unsigned pos = computeChildIdx(key, (int)tree_depth - 1 - (int)depth);
      bool created = false;
      if (!nodeChildExists(node, pos)) {
        if (!nodeHasChildren(node) && !just_created) {
          expandNode(node);
        } else {
          createNodeChild(node, pos);
          created = true;
        }
      }
      node = node->children[pos];
      just_created = created;
    }
    node->value = log_odds_value;
  }

  // finest-level leaves, pruned nodes expanded
  template <class F>
  void forEachLeaf(F f) const {
    if (root) leafRecurs(root, 0, Key(0, 0, 0), f);
  }

  // --- prune / toMaxLikelihood / writeBinary  [octomap OcTreeBaseImpl.hxx, OccupancyOcTreeBase.hxx,
  //     AbstractOccupancyOcTree.cpp] -- behind writeOctomapToFile / writeOctomapToBinaryConst
  //     (octomap_world.cc:849-856)
  void prune() {
    if (root == NULL) return;
    for (unsigned depth = tree_depth - 1; depth > 0; --depth) {
      unsigned num_pruned = 0;
      pruneRecurs(root, 0, depth, num_pruned);
      if (num_pruned == 0) break;
    }
  }
  void toMaxLikelihood() {
    if (root) toMaxLikelihoodRecurs(root, 0, tree_depth);
  }
  void writeBinaryConst(std::string& out) const {
    char buf[256];
    out += "# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
    out += "id OcTree\n";
    snprintf(buf, sizeof(buf), "size %zu\n", tree_size);
    out += buf;
    snprintf(buf, sizeof(buf), "res %g\n", resolution);   // std::ostream << double: 6 significant digits
    out += buf;
    out += "data\n";
    if (root) writeBinaryNode(out, root);
  }
  // OcTreeBaseImpl::write / writeData (full tree: float value + child byte per node)
  void writeFull(std::string& out, bool with_header) const {
    if (with_header) {
      char buf[256];
      out += "# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
      out += "id OcTree\n";
      snprintf(buf, sizeof(buf), "size %zu\n", tree_size);
      out += buf;
      snprintf(buf, sizeof(buf), "res %g\n", resolution);
      out += buf;
      out += "data\n";
    }
    if (root) writeNodesRecurs(out, root);
  }
  void writeBinary(std::string& out) {   // non-const: thresholds and prunes the live tree first
    toMaxLikelihood();
    prune();
    writeBinaryConst(out);
  }

  float prob_hit_log, prob_miss_log, clamping_thres_min, clamping_thres_max, occ_prob_thres_log;

  // change detection (OccupancyOcTreeBase::enableChangeDetection / changedKeysBegin / reset)
  typedef std::unordered_map<Key, bool, KeyHash> KeyBoolMap;
  bool use_change_detection = false;
  KeyBoolMap changed_keys;

 private:
  Node* root;
  size_t tree_size;
  double resolution;
  double resolution_factor;

  void allocNodeChildren(Node* n) {
    n->children = new Node*[8];
    for (unsigned i = 0; i < 8; i++) n->children[i] = NULL;
  }
  Node* createNodeChild(Node* n, unsigned i) {
    if (n->children == NULL) allocNodeChildren(n);
    Node* c = new Node();
    n->children[i] = c;
    tree_size++;
    return c;
  }
  void deleteNodeChild(Node* n, unsigned i) {
    delete n->children[i];
    n->children[i] = NULL;
    tree_size--;
  }
  void deleteNodeRecurs(Node* n) {
    if (n->children != NULL) {
      for (unsigned i = 0; i < 8; i++)
        if (n->children[i] != NULL) deleteNodeRecurs(n->children[i]);
      delete[] n->children;
    }
    delete n;
  }
  void expandNode(Node* n) {
    for (unsigned k = 0; k < 8; k++) {
      Node* c = createNodeChild(n, k);
      c->value = n->value;
    }
  }
  bool isNodeCollapsible(const Node* n) const {
    if (!nodeChildExists(n, 0)) return false;
    const Node* first = n->children[0];
    if (nodeHasChildren(first)) return false;
    for (unsigned i = 1; i < 8; i++) {
      if (!nodeChildExists(n, i) || nodeHasChildren(n->children[i]) ||
          !(n->children[i]->value == first->value))
        return false;
    }
    return true;
  }
  bool pruneNode(Node* n) {
    if (!isNodeCollapsible(n)) return false;
    n->value = n->children[0]->value;
    for (unsigned i = 0; i < 8; i++) deleteNodeChild(n, i);
    delete[] n->children;
    n->children = NULL;
    return true;
  }
  static float getMaxChildLogOdds(const Node* n) {
    float max = -std::numeric_limits<float>::max();
    if (n->children != NULL) {
      for (unsigned i = 0; i < 8; i++) {
        if (n->children[i] != NULL) {
          float l = n->children[i]->value;
          if (l > max) max = l;
        }
      }
    }
    return max;
  }
  void pruneRecurs(Node* node, unsigned depth, unsigned max_depth, unsigned& num_pruned) {
    if (depth < max_depth) {
      for (unsigned i = 0; i < 8; i++)
        if (nodeChildExists(node, i)) pruneRecurs(node->children[i], depth + 1, max_depth, num_pruned);
    } else {
      if (pruneNode(node)) num_pruned++;
    }
  }
  void toMaxLikelihoodRecurs(Node* node, unsigned depth, unsigned max_depth) {
    if (depth < max_depth) {
      for (unsigned i = 0; i < 8; i++)
        if (nodeChildExists(node, i)) toMaxLikelihoodRecurs(node->children[i], depth + 1, max_depth);
    }
    node->value = isNodeOccupied(node) ? clamping_thres_max : clamping_thres_min;   // nodeToMaxLikelihood
  }
  // two bits per child: 10 free, 01 occupied, 00 unknown, 11 has children (bit 2i first)
  void writeBinaryNode(std::string& out, const Node* node) const {
    unsigned char b[2] = {0, 0};
    for (unsigned i = 0; i < 8; i++) {
      if (!nodeChildExists(node, i)) continue;
      const Node* child = node->children[i];
      unsigned two;
      if (nodeHasChildren(child)) two = 3u;
      else if (isNodeOccupied(child)) two = 2u;     // bit[2i] = 0, bit[2i+1] = 1
      else two = 1u;                                // bit[2i] = 1, bit[2i+1] = 0
      b[i >> 2] |= (unsigned char)(two << (2 * (i & 3u)));
    }
    out.push_back((char)b[0]);
    out.push_back((char)b[1]);
    for (unsigned i = 0; i < 8; i++)
      if (nodeChildExists(node, i) && nodeHasChildren(node->children[i])) writeBinaryNode(out, node->children[i]);
  }
  void writeNodesRecurs(std::string& out, const Node* node) const {
    out.append(reinterpret_cast<const char*>(&node->value), sizeof(float));
    unsigned char children = 0;
    for (unsigned i = 0; i < 8; i++)
      if (nodeChildExists(node, i)) children |= (unsigned char)(1u << i);
    out.push_back((char)children);
    for (unsigned i = 0; i < 8; i++)
      if (children & (1u << i)) writeNodesRecurs(out, node->children[i]);
  }
  void updateNodeLogOdds(Node* n, float update) const {
    n->value = n->value + update;
    if (n->value < clamping_thres_min) {
      n->value = clamping_thres_min;
      return;
    }
    if (n->value > clamping_thres_max) n->value = clamping_thres_max;
  }
  Node* updateNodeRecurs(Node* node, bool node_just_created, const Key& key, unsigned depth,
                         float log_odds_update) {
    bool created_node = false;
    if (depth < tree_depth) {
      unsigned pos = computeChildIdx(key, (int)tree_depth - 1 - (int)depth);
      if (!nodeChildExists(node, pos)) {
        if (!nodeHasChildren(node) && !node_just_created) {
          expandNode(node);  // pruned node
        } else {
          createNodeChild(node, pos);
          created_node = true;
        }
      }
      Node* retval =
          updateNodeRecurs(node->children[pos], created_node, key, depth + 1, log_odds_update);
      if (pruneNode(node)) {
        retval = node;
      } else {
        node->value = getMaxChildLogOdds(node);
      }
      return retval;
    } else {
      if (use_change_detection) {
        bool occBefore = isNodeOccupied(node);
        updateNodeLogOdds(node, log_odds_update);
        if (node_just_created) {  // new node
          changed_keys.insert(std::pair<Key, bool>(key, true));
        } else if (occBefore != isNodeOccupied(node)) {  // occupancy changed, track it
          KeyBoolMap::iterator it = changed_keys.find(key);
          if (it == changed_keys.end())
            changed_keys.insert(std::pair<Key, bool>(key, false));
          else if (it->second == false)
            changed_keys.erase(it);
        }
      } else {
        updateNodeLogOdds(node, log_odds_update);
      }
      return node;
    }
  }
  void updateInnerOccupancyRecurs(Node* node, unsigned depth) {
    if (nodeHasChildren(node)) {
      if (depth < tree_depth) {
        for (unsigned i = 0; i < 8; i++)
          if (nodeChildExists(node, i)) updateInnerOccupancyRecurs(node->children[i], depth + 1);
      }
      node->value = getMaxChildLogOdds(node);
    }
  }
  template <class F>
  void leafRecurs(const Node* n, unsigned depth, Key prefix, F& f) const {
    if (depth == tree_depth) {
      f(prefix, n->value);
      return;
    }
    if (!nodeHasChildren(n)) {
      // pruned leaf: expand to the finest level
      unsigned span = 1u << (tree_depth - depth);
      for (unsigned z = 0; z < span; z++)
        for (unsigned y = 0; y < span; y++)
          for (unsigned x = 0; x < span; x++)
            f(Key((key_type)(prefix.k[0] + x), (key_type)(prefix.k[1] + y),
                  (key_type)(prefix.k[2] + z)),
              n->value);
      return;
    }
    unsigned bit = tree_depth - 1 - depth;
    for (unsigned i = 0; i < 8; i++) {
      if (!nodeChildExists(n, i)) continue;
      Key k = prefix;
      if (i & 1) k.k[0] = (key_type)(k.k[0] | (1u << bit));
      if (i & 2) k.k[1] = (key_type)(k.k[1] | (1u << bit));
      if (i & 4) k.k[2] = (key_type)(k.k[2] | (1u << bit));
      leafRecurs(n->children[i], depth + 1, k, f);
    }
  }
};

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Eigen::Quaternion<double>::toRotationMatrix + translation -> row-major 4x4
void quat_to_matrix(const double q[4], const double t[3], double T[16]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T[0] = 1.0 - (tyy + tzz); T[1] = txy - twz;         T[2] = txz + twy;          T[3] = t[0];
  T[4] = txy + twz;         T[5] = 1.0 - (txx + tzz); T[6] = tyz - twx;          T[7] = t[1];
  T[8] = txz - twy;         T[9] = tyz + twx;         T[10] = 1.0 - (txx + tyy); T[11] = t[2];
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// kindr::minimal::QuatTransformation::operator*(Vector3d) = q._transformVector(v) + t
// Eigen: uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
void quat_transform(const double q[4], const double t[3], const double v[3], double out[3]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  double uv[3];
  uv[0] = y * v[2] - z * v[1];
  uv[1] = z * v[0] - x * v[2];
  uv[2] = x * v[1] - y * v[0];
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3];
  c[0] = y * uv[2] - z * uv[1];
  c[1] = z * uv[0] - x * uv[2];
  c[2] = x * uv[1] - y * uv[0];
  for (int i = 0; i < 3; i++) out[i] = ((v[i] + w * uv[i]) + c[i]) + t[i];
}

}  // namespace

// -----------------------------------------------------------------------------------------
// OctomapWorld restatement
// -----------------------------------------------------------------------------------------
struct oracle_world {
  oracle_params params;
  OcTree* octree;
  KeyRay key_ray;  // scratch member, octomap_world.h:326-328
  bool faithful_inner;
  KeySet last_free, last_occ;
  oracle_scan_stats stats;

  oracle_world() : octree(NULL), faithful_inner(true) { std::memset(&stats, 0, sizeof(stats)); }
  ~oracle_world() { delete octree; }

  // octomap_world.cc:84-104
  void set_params(const oracle_params& p) {
    if (octree) {
      if (octree->getResolution() != p.resolution) {
        delete octree;
        octree = new OcTree(p.resolution);
      }
    } else {
      octree = new OcTree(p.resolution);
    }
    octree->setProbHit(p.probability_hit);
    octree->setProbMiss(p.probability_miss);
    octree->setClampingThresMin(p.threshold_min);
    octree->setClampingThresMax(p.threshold_max);
    octree->setOccupancyThres(p.threshold_occupancy);
    octree->use_change_detection = p.change_detection_enabled != 0;   // octomap_world.cc:99
    params = p;
  }

  // statistics only (not in the reference): the range predicate of castRay for EVERY valid
  // point, cast or skipped, so the counter is comparable with the GPU path's.
  void count_in_range(const Vec3f& sensor_origin, const Vec3f& point) {
    if (params.sensor_max_range < 0.0 ||
        (point - sensor_origin).norm() <= params.sensor_max_range)
      stats.points_in_range++;
  }

  // octomap_world.cc:177-230
  void castRay(const Vec3f& sensor_origin, const Vec3f& point, KeySet* free_cells,
               KeySet* occupied_cells) {
    stats.rays_cast++;
    if (params.sensor_max_range < 0.0 ||
        (point - sensor_origin).norm() <= params.sensor_max_range) {
      key_ray.reset();
      if (octree->computeRayKeys(sensor_origin, point, key_ray)) {
        collectFree(sensor_origin, free_cells);
      }
      Key key;
      if (octree->coordToKeyChecked(point, key)) {
        occupied_cells->insert(key);
        stats.occupied_emitted++;
      }
    } else {
      Vec3f new_end =
          sensor_origin + (point - sensor_origin).normalized() * (float)params.sensor_max_range;
      key_ray.reset();
      if (octree->computeRayKeys(sensor_origin, new_end, key_ray)) {
        collectFree(sensor_origin, free_cells);
      }
    }
  }
  // octomap_world.cc:189-201 / 215-227
  void collectFree(const Vec3f& sensor_origin, KeySet* free_cells) {
    const size_t n = key_ray.size();
    stats.traversals += n;
    if (n > stats.longest_ray) stats.longest_ray = n;
    if (params.max_free_space == 0.0) {
      free_cells->insert(key_ray.ray.begin(), key_ray.ray.begin() + n);
    } else {
      for (size_t i = 0; i < n; i++) {
        const Key& key = key_ray.ray[i];
        Vec3f voxel_coordinate = octree->keyToCoord(key);
        if ((voxel_coordinate - sensor_origin).norm() < params.max_free_space ||
            voxel_coordinate(2) > (sensor_origin(2) - params.min_height_free_space)) {
          free_cells->insert(key);
        }
      }
    }
  }

  // octomap_world.cc:238-264
  void updateOccupancy(KeySet* free_cells, KeySet* occupied_cells) {
    double t0 = now_s();
    for (KeySet::iterator it = occupied_cells->begin(), end = occupied_cells->end(); it != end;
         it++) {
      octree->updateNode(*it, true);
      if (free_cells->find(*it) != free_cells->end()) free_cells->erase(*it);
    }
    for (KeySet::iterator it = free_cells->begin(), end = free_cells->end(); it != end; ++it) {
      octree->updateNode(*it, false);
    }
    double t1 = now_s();
    if (faithful_inner) octree->updateInnerOccupancy();
    double t2 = now_s();
    stats.t_update_s = t1 - t0;
    stats.t_inner_s = t2 - t1;
    stats.unique_free = free_cells->size();
    stats.unique_occupied = occupied_cells->size();
  }

  void begin_scan(uint64_t points_in) {
    std::memset(&stats, 0, sizeof(stats));
    stats.points_in = points_in;
    last_free.clear();
    last_occ.clear();
  }
};

extern "C" {

void oracle_default_params(oracle_params* p) {
  // octomap_world.h:56-69
  std::memset(p, 0, sizeof(*p));
  p->resolution = 0.15;
  p->probability_hit = 0.65;
  p->probability_miss = 0.4;
  p->threshold_min = 0.12;
  p->threshold_max = 0.97;
  p->threshold_occupancy = 0.7;
  p->filter_speckles = 1;
  p->max_free_space = 0.0;
  p->min_height_free_space = 0.0;
  p->sensor_max_range = 5.0;
  p->visualize_min_z = -std::numeric_limits<double>::max();
  p->visualize_max_z = std::numeric_limits<double>::max();
  p->treat_unknown_as_occupied = 1;
  p->change_detection_enabled = 0;
}

oracle_world* oracle_create(const oracle_params* p) {
  oracle_world* w = new oracle_world();
  oracle_params d;
  if (!p) { oracle_default_params(&d); p = &d; }
  w->set_params(*p);
  return w;
}
void oracle_destroy(oracle_world* w) { delete w; }
void oracle_reset(oracle_world* w) { w->octree->clear(); }  // octomap_world.cc:75-80
void oracle_set_params(oracle_world* w, const oracle_params* p) { w->set_params(*p); }
void oracle_get_params(const oracle_world* w, oracle_params* p) { *p = w->params; }
void oracle_set_faithful_inner_update(oracle_world* w, int on) { w->faithful_inner = on != 0; }

// octomap_world.cc:110-141
void oracle_insert_pointcloud(oracle_world* w, float* xyz, size_t n, size_t stride_bytes,
                              int is_dense, const double T[16], size_t* n_out) {
  w->begin_scan(n);
  double t0 = now_s();
  char* base = reinterpret_cast<char*>(xyz);
  auto P = [&](size_t i) { return reinterpret_cast<float*>(base + i * stride_bytes); };

  // pcl::removeNaNFromPointCloud(*cloud, *cloud, indices): no-op when is_dense, else stable
  // compaction of points whose three coordinates are all finite.
  size_t m = n;
  if (!is_dense) {
    size_t j = 0;
    for (size_t i = 0; i < n; i++) {
      float* p = P(i);
      if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
      if (j != i) std::memmove(P(j), p, 3 * sizeof(float));
      j++;
    }
    m = j;
  }
  w->stats.points_valid = m;
  if (n_out) *n_out = m;

  // pcl::transformPointCloud(*cloud, *cloud, Matrix4d) -- dense branch, double maths, scalar
  // left-to-right association, float store.
  for (size_t i = 0; i < m; i++) {
    float* p = P(i);
    double x = p[0], y = p[1], z = p[2];
    float ox = static_cast<float>(T[0] * x + T[1] * y + T[2] * z + T[3]);
    float oy = static_cast<float>(T[4] * x + T[5] * y + T[6] * z + T[7]);
    float oz = static_cast<float>(T[8] * x + T[9] * y + T[10] * z + T[11]);
    p[0] = ox; p[1] = oy; p[2] = oz;
  }
  // pointEigenToOctomap(T_G_sensor.getPosition())
  const Vec3f p_G_sensor((float)T[3], (float)T[7], (float)T[11]);

  KeySet& free_cells = w->last_free;
  KeySet& occupied_cells = w->last_occ;
  for (size_t i = 0; i < m; i++) {
    const float* p = P(i);
    const Vec3f p_G_point(p[0], p[1], p[2]);
    Key key = w->octree->coordToKey(p_G_point);
    w->count_in_range(p_G_sensor, p_G_point);
    if (occupied_cells.find(key) == occupied_cells.end()) {
      w->castRay(p_G_sensor, p_G_point, &free_cells, &occupied_cells);
    }
  }
  double t1 = now_s();
  w->updateOccupancy(&free_cells, &occupied_cells);
  w->stats.t_keyset_s = t1 - t0;
}

// octomap_world.cc:143-175
void oracle_insert_projected(oracle_world* w, const float* xyz, int rows, int cols,
                             size_t row_pitch_bytes, const double q[4], const double t[3]) {
  w->begin_scan((uint64_t)rows * (uint64_t)cols);
  double t0 = now_s();
  const double zero[3] = {0.0, 0.0, 0.0};
  double o[3];
  quat_transform(q, t, zero, o);
  const Vec3f sensor_origin((float)o[0], (float)o[1], (float)o[2]);

  KeySet& free_cells = w->last_free;
  KeySet& occupied_cells = w->last_occ;
  for (int v = 0; v < rows; ++v) {
    const float* row = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xyz) +
                                                      (size_t)v * row_pitch_bytes);
    for (int u = 0; u < cols; ++u) {
      const float* pt = row + 3 * u;
      // isValidPoint (octomap_world.cc:232-236) || z < 0
      const bool valid = pt[2] != 10000.0f && !std::isinf(pt[2]);
      if (!valid || pt[2] < 0) continue;
      w->stats.points_valid++;
      const double pe[3] = {(double)pt[0], (double)pt[1], (double)pt[2]};
      double pw[3];
      quat_transform(q, t, pe, pw);
      const Vec3f point_octomap((float)pw[0], (float)pw[1], (float)pw[2]);
      Key key = w->octree->coordToKey(point_octomap);
      w->count_in_range(sensor_origin, point_octomap);
      if (occupied_cells.find(key) == occupied_cells.end()) {
        w->castRay(sensor_origin, point_octomap, &free_cells, &occupied_cells);
      }
    }
  }
  double t1 = now_s();
  w->updateOccupancy(&free_cells, &occupied_cells);
  w->stats.t_keyset_s = t1 - t0;
}

// world_base.cc:55-69
void oracle_fixup_q(const double Q_full[16], double full_image_width, int disparity_cols,
                    double Q[16]) {
  std::memcpy(Q, Q_full, 16 * sizeof(double));
  double downsampling_factor = full_image_width / disparity_cols;
  if (std::fabs(downsampling_factor - 1.0) > 1e-6) {
    const double downsampling_squared = downsampling_factor * downsampling_factor;
    Q[0 * 4 + 0] /= downsampling_factor;
    Q[0 * 4 + 3] /= downsampling_squared;
    Q[1 * 4 + 1] /= downsampling_factor;
    Q[1 * 4 + 3] /= downsampling_squared;
    Q[2 * 4 + 3] /= downsampling_squared;
    Q[3 * 4 + 2] /= downsampling_factor;
    Q[3 * 4 + 3] /= downsampling_factor;
  }
}

// cv::reprojectImageTo3D(disparity CV_32F, out CV_32FC3, Q, handleMissingValues=true), 4.13
void oracle_reproject(const float* disp, int rows, int cols, size_t row_pitch_bytes,
                      const double Q[16], float* out) {
  const float bigZ = 10000.f;
  double minDisparity = FLT_MAX;
  // cv::minMaxIdx over the whole image
  bool first = true;
  for (int y = 0; y < rows; y++) {
    const float* s = reinterpret_cast<const float*>(reinterpret_cast<const char*>(disp) +
                                                    (size_t)y * row_pitch_bytes);
    for (int x = 0; x < cols; x++) {
      if (first || (double)s[x] < minDisparity) { minDisparity = s[x]; first = false; }
    }
  }
  for (int y = 0; y < rows; y++) {
    const float* s = reinterpret_cast<const float*>(reinterpret_cast<const char*>(disp) +
                                                    (size_t)y * row_pitch_bytes);
    float* dptr = out + (size_t)y * cols * 3;
    for (int x = 0; x < cols; x++) {
      double d = s[x];
      double h[4];
      for (int r = 0; r < 4; r++)
        h[r] = ((Q[r * 4 + 0] * (double)x + Q[r * 4 + 1] * (double)y) + Q[r * 4 + 2] * d) +
               Q[r * 4 + 3] * 1.0;
      const double iw = 1.0 / h[3];
      for (int r = 0; r < 3; r++) dptr[3 * x + r] = (float)((double)(float)h[r] * iw);
      if (std::fabs(d - minDisparity) <= FLT_EPSILON) dptr[3 * x + 2] = bigZ;
    }
  }
}

void oracle_insert_disparity(oracle_world* w, const float* disp, int rows, int cols,
                             size_t row_pitch_bytes, const double Q_full[16],
                             double full_image_width, const double q[4], const double t[3]) {
  double Q[16];
  oracle_fixup_q(Q_full, full_image_width, cols, Q);
  std::vector<float> xyz((size_t)rows * cols * 3);
  oracle_reproject(disp, rows, cols, row_pitch_bytes, Q, xyz.data());
  oracle_insert_projected(w, xyz.data(), rows, cols, (size_t)cols * 3 * sizeof(float), q, t);
}

void oracle_apply_keys(oracle_world* w, const uint16_t* free_k3, size_t n_free,
                       const uint16_t* occ_k3, size_t n_occ) {
  w->begin_scan(0);
  for (size_t i = 0; i < n_free; i++)
    w->last_free.insert(Key(free_k3[3 * i], free_k3[3 * i + 1], free_k3[3 * i + 2]));
  for (size_t i = 0; i < n_occ; i++)
    w->last_occ.insert(Key(occ_k3[3 * i], occ_k3[3 * i + 1], occ_k3[3 * i + 2]));
  w->updateOccupancy(&w->last_free, &w->last_occ);
}

// octomap_world.cc:346-360
void oracle_query_points(const oracle_world* w, const double* xyz, size_t n, uint8_t* status) {
  for (size_t i = 0; i < n; i++) {
    Node* node = w->octree->search(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (node == NULL) {
      status[i] = w->params.treat_unknown_as_occupied ? 1 : 2;
    } else if (w->octree->isNodeOccupied(node)) {
      status[i] = 1;
    } else {
      status[i] = 0;
    }
  }
}

// octomap_world.cc:395-418
void oracle_query_lines(const oracle_world* w, const double* ab, size_t n, uint8_t* status) {
  for (size_t i = 0; i < n; i++) {
    KeyRay key_ray;  // constructed per call, as the reference does (:400)
    const Vec3f a((float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2]);
    const Vec3f b((float)ab[6 * i + 3], (float)ab[6 * i + 4], (float)ab[6 * i + 5]);
    w->octree->computeRayKeys(a, b, key_ray);
    uint8_t st = 0;
    for (size_t j = 0; j < key_ray.size(); j++) {
      Node* node = w->octree->search(key_ray.ray[j]);
      if (node == NULL) {
        st = w->params.treat_unknown_as_occupied ? 1 : 2;
        break;
      } else if (w->octree->isNodeOccupied(node)) {
        st = 1;
        break;
      }
    }
    status[i] = st;
  }
}

// octomap_world.cc:1413-1440: getChangedPoints -- keys + current occupancy, then reset.
size_t oracle_get_changed(oracle_world* w, uint16_t* k3, uint8_t* occupied, size_t cap) {
  OcTree* t = w->octree;
  const size_t n = t->changed_keys.size();
  if (!k3 || !occupied || cap < n) return n;
  size_t i = 0;
  for (OcTree::KeyBoolMap::const_iterator it = t->changed_keys.begin(); it != t->changed_keys.end(); ++it, ++i) {
    Node* node = t->search(it->first);
    k3[3 * i] = it->first.k[0]; k3[3 * i + 1] = it->first.k[1]; k3[3 * i + 2] = it->first.k[2];
    occupied[i] = (node != NULL && t->isNodeOccupied(node)) ? 1 : 0;
  }
  t->changed_keys.clear();   // resetChangeDetection()
  return n;
}

// octomap_world.cc:781-818 (setLogOddsBoundingBox) with adjustBoundingBox (:1175-1224)
void oracle_set_log_odds_bbox(oracle_world* w, const double* positions, size_t n, const double bbox[3],
                              double log_odds_value, int insertion_method) {
  OcTree* t = w->octree;
  const double resolution = t->getResolution();
  const double epsilon = 0.001;
  for (size_t p = 0; p < n; p++) {
    double bbx_min[3], bbx_max[3];
    for (int a = 0; a < 3; a++) {
      const double position = positions[3 * p + a], size = bbox[a];
      if (insertion_method == 0) {
        bbx_min[a] = position - size / 2 - epsilon;
        bbx_max[a] = position + size / 2 + epsilon;
      } else if (insertion_method == 1) {
        bbx_min[a] = position - size / 2 + epsilon;
        bbx_max[a] = position + size / 2 - epsilon;
        bbx_min[a] = (double)(float)t->keyToCoord(t->coordToKey((double)(float)bbx_min[a]));
        bbx_max[a] = (double)(float)t->keyToCoord(t->coordToKey((double)(float)bbx_max[a]));
        bbx_min[a] -= epsilon;
        bbx_max[a] += epsilon;
      } else {
        bbx_min[a] = position - size / 2 - epsilon;
        bbx_max[a] = position + size / 2 + epsilon;
        bbx_min[a] = (double)(float)t->keyToCoord(t->coordToKey((double)(float)bbx_min[a]));
        bbx_max[a] = (double)(float)t->keyToCoord(t->coordToKey((double)(float)bbx_max[a]));
        bbx_min[a] += resolution;
        bbx_max[a] -= resolution;
        bbx_min[a] -= epsilon;
        bbx_max[a] += epsilon;
      }
    }
    for (double x = bbx_min[0]; x <= bbx_max[0]; x += resolution)
      for (double y = bbx_min[1]; y <= bbx_max[1]; y += resolution)
        for (double z = bbx_min[2]; z <= bbx_max[2]; z += resolution)
          t->setNodeValueLazy(Vec3f((float)x, (float)y, (float)z), (float)log_odds_value);
  }
  t->updateInnerOccupancy();
}

// octomap_world.cc:363-393: getCellTrueStatusPoint + getCellProbabilityPoint
void oracle_query_points_probability(const oracle_world* w, const double* xyz, size_t n, uint8_t* status,
                                     double* probability) {
  for (size_t i = 0; i < n; i++) {
    Node* node = w->octree->search(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (node == NULL) {
      if (probability) probability[i] = -1.0;
      status[i] = 2;
    } else {
      // OcTreeNode::getOccupancy() = probability(value) = 1. - ( 1. / (1. + exp(logodds)))
      if (probability) probability[i] = 1. - (1. / (1. + std::exp((double)node->value)));
      status[i] = w->octree->isNodeOccupied(node) ? 1 : 0;
    }
  }
}

// octomap_world.cc:858-880.  As written upstream the loop searches `key` itself (not
// `current_key`) 26 times, so an occupied node always "finds a neighbour" and is never a speckle;
// restated as is.
static bool is_speckle_node(const OcTree* t, const Key& key) {
  for (int n = 0; n < 26; n++) {
    Node* node = t->search(key);
    if (node && t->isNodeOccupied(node)) return false;
  }
  return true;
}

// octomap_world.cc:274-344 getCellStatusBoundingBox
static uint8_t bbox_status(const oracle_world* w, const double point[3], const double size[3]) {
  const OcTree* t = w->octree;
  const uint8_t unknown_as = w->params.treat_unknown_as_occupied ? 1 : 2;
  // center point unknown or occupied: quit
  Node* cn = t->search(point[0], point[1], point[2]);
  const uint8_t center_status = cn == NULL ? unknown_as : (t->isNodeOccupied(cn) ? 1 : 0);
  if (center_status != 0) return center_status;
  Key key;
  if (!t->coordToKeyChecked(Vec3f((float)point[0], (float)point[1], (float)point[2]), key)) return unknown_as;
  double mn[3], mx[3];
  for (int a = 0; a < 3; a++) {
    mn[a] = point[a] - size[a] / 2;
    mx[a] = point[a] + size[a] / 2;
  }
  const Vec3f bbx_min((float)mn[0], (float)mn[1], (float)mn[2]), bbx_max((float)mx[0], (float)mx[1], (float)mx[2]);
  bool occupied = false;
  t->forEachLeafBBX(bbx_min, bbx_max, [&](const Node* node, const Key& k, unsigned depth) {
    const double cube_size = t->getNodeSize(depth);
    // "Check if it is really inside bounding box, since leaf_bbx_iterator begins too early"
    for (int a = 0; a < 3; a++) {
      const double c = t->keyToCoord(k.k[a], depth);
      const double lower = c - (cube_size / 2) * 1.0, upper = c + (cube_size / 2) * 1.0;
      if (upper < (double)bbx_min(a) || lower > (double)bbx_max(a)) return true;   // continue
    }
    if (t->isNodeOccupied(node)) {
      if (w->params.filter_speckles && is_speckle_node(t, k)) return true;
      occupied = true;
      return false;
    }
    return true;
  });
  if (occupied) return 1;
  if (t->anyUnknownLeafCenter(bbx_min, bbx_max)) return unknown_as;
  return 0;
}

void oracle_query_bbox_status(const oracle_world* w, const double* xyz, size_t n, const double bounding_box_size[3],
                              uint8_t* status) {
  for (size_t i = 0; i < n; i++) status[i] = bbox_status(w, xyz + 3 * i, bounding_box_size);
}

// octomap_world.cc:1384-1412 checkPathForCollisionsWithRobot / checkSinglePoseCollision: returns
// 1 and the index of the earliest colliding pose, or 0.
int oracle_check_path_collisions(const oracle_world* w, const double* xyz, size_t n, const double robot_size[3],
                                 size_t* collision_index) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t st = bbox_status(w, xyz + 3 * i, robot_size);
    const bool hit = w->params.treat_unknown_as_occupied ? (st != 0) : (st == 1);
    if (hit) {
      if (collision_index) *collision_index = i;
      return 1;
    }
  }
  return 0;
}

// octomap_world.cc:420-449
void oracle_query_visibility(const oracle_world* w, const double* ab, size_t n, int stop_at_unknown_cell,
                             uint8_t* status) {
  for (size_t i = 0; i < n; i++) {
    KeyRay key_ray;
    const Vec3f a((float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2]);
    const Vec3f b((float)ab[6 * i + 3], (float)ab[6 * i + 4], (float)ab[6 * i + 5]);
    w->octree->computeRayKeys(a, b, key_ray);
    const Key voxel_to_test_key = w->octree->coordToKey(b);
    uint8_t st = 0;
    for (size_t j = 0; j < key_ray.size(); j++) {
      const Key& key = key_ray.ray[j];
      if (key == voxel_to_test_key) continue;
      Node* node = w->octree->search(key);
      if (node == NULL) {
        if (stop_at_unknown_cell) { st = 2; break; }
      } else if (w->octree->isNodeOccupied(node)) {
        st = 1;
        break;
      }
    }
    status[i] = st;
  }
}

// octomap_world.cc:451-493
void oracle_query_lines_bbox(const oracle_world* w, const double* ab, size_t n, const double bounding_box_size[3],
                             uint8_t* status) {
  const double epsilon = 0.001;
  const double resolution = w->octree->getResolution();
  double x_disc = bounding_box_size[0] / ceil((bounding_box_size[0] + epsilon) / resolution);
  double y_disc = bounding_box_size[1] / ceil((bounding_box_size[1] + epsilon) / resolution);
  double z_disc = bounding_box_size[2] / ceil((bounding_box_size[2] + epsilon) / resolution);
  if (x_disc <= 0.0) x_disc = 1.0;
  if (y_disc <= 0.0) y_disc = 1.0;
  if (z_disc <= 0.0) z_disc = 1.0;
  const double hx = bounding_box_size[0] * 0.5, hy = bounding_box_size[1] * 0.5, hz = bounding_box_size[2] * 0.5;
  for (size_t i = 0; i < n; i++) {
    uint8_t ret = 0;
    for (double x = -hx; x <= hx && ret == 0; x += x_disc)
      for (double y = -hy; y <= hy && ret == 0; y += y_disc)
        for (double z = -hz; z <= hz && ret == 0; z += z_disc) {
          const double seg[6] = {ab[6 * i] + x,     ab[6 * i + 1] + y, ab[6 * i + 2] + z,
                                 ab[6 * i + 3] + x, ab[6 * i + 4] + y, ab[6 * i + 5] + z};
          oracle_query_lines(w, seg, 1, &ret);
        }
    status[i] = ret;
  }
}

int oracle_coord_to_key_checked(const oracle_world* w, double c, uint16_t* key) {
  key_type k = 0;
  bool ok = w->octree->coordToKeyChecked(c, k);
  if (ok && key) *key = k;
  return ok ? 1 : 0;
}
uint16_t oracle_coord_to_key(const oracle_world* w, double c) { return w->octree->coordToKey(c); }
double oracle_key_to_coord(const oracle_world* w, uint16_t k) { return w->octree->keyToCoord(k); }

long oracle_compute_ray_keys(const oracle_world* w, const float origin[3], const float end[3],
                             uint16_t* out_k3, size_t cap) {
  KeyRay ray;
  bool ok = w->octree->computeRayKeys(Vec3f(origin[0], origin[1], origin[2]),
                                      Vec3f(end[0], end[1], end[2]), ray);
  if (!ok) return -1;
  for (size_t i = 0; i < ray.size() && i < cap; i++) {
    out_k3[3 * i] = ray.ray[i].k[0];
    out_k3[3 * i + 1] = ray.ray[i].k[1];
    out_k3[3 * i + 2] = ray.ray[i].k[2];
  }
  return (long)ray.size();
}

void oracle_quat_to_matrix(const double q[4], const double t[3], double T[16]) {
  quat_to_matrix(q, t, T);
}

void oracle_logodds_constants(const oracle_world* w, float out5[5]) {
  out5[0] = w->octree->prob_hit_log;
  out5[1] = w->octree->prob_miss_log;
  out5[2] = w->octree->clamping_thres_min;
  out5[3] = w->octree->clamping_thres_max;
  out5[4] = w->octree->occ_prob_thres_log;
}

size_t oracle_last_free_count(const oracle_world* w) { return w->last_free.size(); }
size_t oracle_last_occupied_count(const oracle_world* w) { return w->last_occ.size(); }
static void dump_keys(const KeySet& s, uint16_t* out) {
  size_t i = 0;
  for (KeySet::const_iterator it = s.begin(); it != s.end(); ++it, ++i) {
    out[3 * i] = it->k[0];
    out[3 * i + 1] = it->k[1];
    out[3 * i + 2] = it->k[2];
  }
}
void oracle_last_free_keys(const oracle_world* w, uint16_t* out) { dump_keys(w->last_free, out); }
void oracle_last_occupied_keys(const oracle_world* w, uint16_t* out) { dump_keys(w->last_occ, out); }
void oracle_last_scan_stats(const oracle_world* w, oracle_scan_stats* s) { *s = w->stats; }

size_t oracle_leaf_count(const oracle_world* w) {
  size_t n = 0;
  w->octree->forEachLeaf([&](const Key&, float) { n++; });
  return n;
}
void oracle_export_leaves(const oracle_world* w, uint16_t* out_k3, float* out_logodds) {
  size_t i = 0;
  w->octree->forEachLeaf([&](const Key& k, float v) {
    out_k3[3 * i] = k.k[0];
    out_k3[3 * i + 1] = k.k[1];
    out_k3[3 * i + 2] = k.k[2];
    out_logodds[i] = v;
    i++;
  });
}
size_t oracle_tree_node_count(const oracle_world* w) { return w->octree->size(); }

void oracle_prune(oracle_world* w) { w->octree->prune(); }

// getOctomapFullMsg payload (with_header == 0) or a complete .ot stream; returns the size.
size_t oracle_write_ot(const oracle_world* w, int with_header, unsigned char* buf, size_t cap) {
  std::string out;
  w->octree->writeFull(out, with_header != 0);
  if (buf && cap >= out.size()) std::memcpy(buf, out.data(), out.size());
  return out.size();
}

// writeOctomapToFile (live != 0: OcTree::writeBinary thresholds + prunes the live tree) or
// writeOctomapToBinaryConst (live == 0) into a caller buffer; returns the stream size.
size_t oracle_write_bt(oracle_world* w, int live, unsigned char* buf, size_t cap) {
  std::string out;
  if (live) w->octree->writeBinary(out);
  else w->octree->writeBinaryConst(out);
  if (buf && cap >= out.size()) std::memcpy(buf, out.data(), out.size());
  return out.size();
}

}  // extern "C"
