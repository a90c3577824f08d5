// vmap_export_host.inl -- on-demand export of the device map to octomap's tree form, on the host
// (included by vmap.cu).  BASELINE.json north_star: "Export to octomap::OcTree, with pruning and
// inner-node max, happens on demand on the host for the unchanged publish/save paths"
// (octomap_world.cc:820-856: get*Msg, setOctomapFromMsg, loadOctomapFromFile, writeOctomapToFile,
// writeOctomapToBinaryConst).
//
// The finest-level leaves are read back, put into Morton order (== octree depth-first order) and
// folded bottom-up into octomap's canonical pruned tree: a node whose 8 children exist, are
// leaves and carry equal values collapses into one leaf; every other inner node takes the maximum
// of its children (OccupancyOcTreeBase::prune / updateInnerOccupancy).  The ".bt"
// stream follows OcTreeBase::writeBinary: text header, then depth-first two bytes per inner node,
// two bits per child (00 none, 01 occupied leaf, 10 free leaf, 11 inner node).  octomap itself is
// not available here, so the byte layout is restated from the upstream sources [UPSTREAM-recalled]
// and pinned by a hand-derived stream and by the oracle's independent pointer-tree writer.
#include <fstream>
#include <memory>
#include <sstream>

namespace {

struct TreeNode {
  uint64_t code;    // Morton prefix: (depth * 3) significant bits
  float value;      // leaf log-odds, or max of the children for an inner node
  uint8_t inner;    // 1: has children
  uint8_t alive;    // 0: merged into its parent by pruning
};

struct PrunedTree {
  std::vector<TreeNode> level[17];   // level[d]: nodes at depth d, Morton order; level[16] = voxels
  size_t n_nodes = 0;
};

uint64_t host_spread3(uint32_t v) { return spread3(v); }

int read_leaves_morton(vmap* h, std::vector<std::pair<uint64_t, float> >* out) {
  size_t n = 0;
  VM_TRY(vmap_leaf_count(h, &n));
  std::vector<uint16_t> k3(n * 3);
  std::vector<float> val(n);
  if (n) VM_TRY(vmap_export_leaves(h, k3.data(), val.data(), n, &n));
  out->resize(n);
  for (size_t i = 0; i < n; i++)
    (*out)[i] = std::make_pair(host_spread3(k3[3 * i]) | (host_spread3(k3[3 * i + 1]) << 1) | (host_spread3(k3[3 * i + 2]) << 2), val[i]);
  std::sort(out->begin(), out->end(), [](const std::pair<uint64_t, float>& a, const std::pair<uint64_t, float>& b) { return a.first < b.first; });
  return VMAP_OK;
}

// max_likelihood: leaves are first thresholded to the clamping values (toMaxLikelihood).
void build_pruned_tree(const vmap* h, const std::vector<std::pair<uint64_t, float> >& leaves, bool max_likelihood,
                       PrunedTree* t) {
  for (auto& l : t->level) l.clear();
  t->n_nodes = 0;
  std::vector<TreeNode>& fin = t->level[16];
  fin.reserve(leaves.size());
  for (const auto& kv : leaves) {
    float v = kv.second;
    if (max_likelihood) v = (v >= h->P.occ_thres) ? h->P.cmax : h->P.cmin;
    fin.push_back(TreeNode{kv.first, v, 0, 1});
  }
  for (int d = 15; d >= 0; d--) {
    std::vector<TreeNode>& child = t->level[d + 1];
    std::vector<TreeNode>& par = t->level[d];
    size_t i = 0;
    while (i < child.size()) {
      const uint64_t pc = child[i].code >> 3;
      size_t j = i;
      bool all_leaf = true, equal = true;
      float mx = -std::numeric_limits<float>::max();
      while (j < child.size() && (child[j].code >> 3) == pc) {
        if (child[j].inner) all_leaf = false;
        if (!(child[j].value == child[i].value)) equal = false;
        if (child[j].value > mx) mx = child[j].value;
        j++;
      }
      if (d > 0 && j - i == 8 && all_leaf && equal) {   // prune() never collapses into the root
        for (size_t k = i; k < j; k++) child[k].alive = 0;
        par.push_back(TreeNode{pc, child[i].value, 0, 1});
      } else {
        par.push_back(TreeNode{pc, mx, 1, 1});
      }
      i = j;
    }
  }
  for (int d = 0; d <= 16; d++)
    for (const TreeNode& n : t->level[d]) t->n_nodes += n.alive;
}

size_t first_child(const PrunedTree& t, int d, uint64_t code) {
  const std::vector<TreeNode>& c = t.level[d + 1];
  size_t lo = 0, hi = c.size();
  const uint64_t want = code << 3;
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (c[mid].code < want) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// OcTreeBase::writeBinaryNode, iteratively (explicit stack: depth <= 16)
void write_bt_nodes(const vmap* h, const PrunedTree& t, std::string* out) {
  if (t.level[0].empty()) return;
  struct Frame { int d; uint64_t code; };
  std::vector<Frame> stack;
  if (!t.level[0][0].inner) return;   // a single pruned root leaf has no child record
  stack.push_back(Frame{0, t.level[0][0].code});
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    const std::vector<TreeNode>& c = t.level[f.d + 1];
    size_t i = first_child(t, f.d, f.code);
    uint32_t bits = 0;
    Frame kids[8];
    int n_kids = 0;
    for (; i < c.size() && (c[i].code >> 3) == f.code; i++) {
      const uint32_t pos = (uint32_t)(c[i].code & 7u);
      uint32_t two;
      if (c[i].inner) {
        two = 3u;   // 11: child has children
        kids[n_kids++] = Frame{f.d + 1, c[i].code};
      } else if (c[i].value >= h->P.occ_thres) {
        two = 2u;   // bit(2i)=0, bit(2i+1)=1: occupied
      } else {
        two = 1u;   // bit(2i)=1, bit(2i+1)=0: free
      }
      bits |= two << (2 * pos);
    }
    out->push_back((char)(bits & 0xFFu));          // child1to4
    out->push_back((char)((bits >> 8) & 0xFFu));   // child5to8
    for (int k = n_kids - 1; k >= 0; k--) stack.push_back(kids[k]);   // depth-first, child 0 first
  }
}

// OcTreeBaseImpl::writeNodesRecurs (the full ".ot" form behind getOctomapFullMsg /
// fullMapToMsg, octomap_world.cc:826-831): per node its float value, then one byte with one bit
// per existing child, then the children depth-first.
void write_ot_nodes(const PrunedTree& t, std::string* out) {
  if (t.level[0].empty()) return;
  struct Frame { int d; size_t idx; };
  std::vector<Frame> stack;
  stack.push_back(Frame{0, 0});
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    const TreeNode& n = t.level[f.d][f.idx];
    out->append(reinterpret_cast<const char*>(&n.value), sizeof(float));
    uint32_t bits = 0;
    Frame kids[8];
    int n_kids = 0;
    if (n.inner) {
      const std::vector<TreeNode>& c = t.level[f.d + 1];
      for (size_t i = first_child(t, f.d, n.code); i < c.size() && (c[i].code >> 3) == n.code; i++) {
        bits |= 1u << (uint32_t)(c[i].code & 7u);
        kids[n_kids++] = Frame{f.d + 1, i};
      }
    }
    out->push_back((char)bits);
    for (int k = n_kids - 1; k >= 0; k--) stack.push_back(kids[k]);
  }
}

std::string ot_header(const vmap* h, size_t n_nodes) {
  std::ostringstream s;
  s << "# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
  s << "id OcTree" << std::endl;
  s << "size " << n_nodes << std::endl;
  s << "res " << h->params.resolution << std::endl;
  s << "data" << std::endl;
  return s.str();
}

// OcTreeBaseImpl::readNodesRecurs: leaves keep their float values; pruned leaves are expanded.
int parse_ot(vmap* h, const std::string& in, bool with_header, double* res_out, std::vector<uint16_t>* k3,
             std::vector<float>* val) {
  size_t pos = 0;
  *res_out = h->params.resolution;
  if (with_header) {
    auto getline_ = [&](std::string* line) {
      if (pos >= in.size()) return false;
      const size_t e = in.find('\n', pos);
      *line = in.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
      pos = (e == std::string::npos) ? in.size() : e + 1;
      return true;
    };
    std::string line, id;
    if (!getline_(&line) || line.compare(0, 21, "# Octomap OcTree file") != 0)
      return fail(h, VMAP_ERR_INVALID_ARGUMENT, "not an octomap (.ot) stream");
    double res = 0.0;
    bool got_data = false;
    while (getline_(&line)) {
      if (line.empty() || line[0] == '#') continue;
      std::istringstream ls(line);
      std::string tok;
      ls >> tok;
      if (tok == "id") ls >> id;
      else if (tok == "res") ls >> res;
      else if (tok == "data") { got_data = true; break; }
    }
    if (!got_data || id != "OcTree" || !(res > 0.0)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "malformed .ot header");
    *res_out = res;
  }
  k3->clear();
  val->clear();
  if (pos >= in.size()) return VMAP_OK;
  struct Frame { int d; uint32_t x, y, z; };
  std::vector<Frame> stack;
  stack.push_back(Frame{0, 0, 0, 0});
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    if (pos + 5 > in.size()) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "truncated .ot data");
    float v;
    std::memcpy(&v, in.data() + pos, 4);
    const uint32_t bits = (uint8_t)in[pos + 4];
    pos += 5;
    if (bits == 0) {   // leaf at depth f.d
      const uint32_t span = 1u << (16 - f.d);
      if ((double)span * span * span > 16777216.0 || k3->size() / 3 + (size_t)span * span * span > (1ull << 31))
        return fail(h, VMAP_ERR_CAPACITY, "pruned leaf too coarse to expand into finest-level voxels");
      for (uint32_t z = 0; z < span; z++)
        for (uint32_t y = 0; y < span; y++)
          for (uint32_t x = 0; x < span; x++) {
            k3->push_back((uint16_t)(f.x + x)); k3->push_back((uint16_t)(f.y + y)); k3->push_back((uint16_t)(f.z + z));
            val->push_back(v);
          }
      continue;
    }
    if (f.d >= 16) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "inner node below the finest level in .ot data");
    const uint32_t bit = 1u << (15 - f.d);
    Frame kids[8];
    int n_kids = 0;
    for (uint32_t i = 0; i < 8; i++)
      if (bits & (1u << i))
        kids[n_kids++] = Frame{f.d + 1, f.x | ((i & 1u) ? bit : 0u), f.y | ((i & 2u) ? bit : 0u), f.z | ((i & 4u) ? bit : 0u)};
    for (int k = n_kids - 1; k >= 0; k--) stack.push_back(kids[k]);
  }
  return VMAP_OK;
}

std::string bt_header(const vmap* h, size_t n_nodes) {
  std::ostringstream s;
  s << "# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
  s << "id OcTree" << std::endl;
  s << "size " << n_nodes << std::endl;
  s << "res " << h->params.resolution << std::endl;
  s << "data" << std::endl;
  return s.str();
}

int serialize_bt(vmap* h, std::string* out, size_t* n_nodes) {
  std::vector<std::pair<uint64_t, float> > leaves;
  VM_TRY(read_leaves_morton(h, &leaves));
  PrunedTree t;
  build_pruned_tree(h, leaves, true, &t);
  *out = bt_header(h, t.n_nodes);
  write_bt_nodes(h, t, out);
  if (n_nodes) *n_nodes = t.n_nodes;
  return VMAP_OK;
}

// OcTreeBase::readBinaryNode: leaves carry the clamping values; pruned leaves are expanded to
// the finest level (bounded: a coarse leaf would be billions of voxels).
int parse_bt(vmap* h, const std::string& in, double* res_out, std::vector<uint16_t>* k3, std::vector<float>* val) {
  size_t pos = 0;
  auto getline_ = [&](std::string* line) {
    if (pos >= in.size()) return false;
    const size_t e = in.find('\n', pos);
    *line = in.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
    pos = (e == std::string::npos) ? in.size() : e + 1;
    return true;
  };
  std::string line;
  if (!getline_(&line) || line.compare(0, 28, "# Octomap OcTree binary file") != 0)
    return fail(h, VMAP_ERR_INVALID_ARGUMENT, "not an octomap binary (.bt) stream");
  double res = 0.0;
  size_t size = 0;
  bool got_data = false;
  std::string id;
  while (getline_(&line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string tok;
    ls >> tok;
    if (tok == "id") ls >> id;
    else if (tok == "size") ls >> size;
    else if (tok == "res") ls >> res;
    else if (tok == "data") { got_data = true; break; }
  }
  if (!got_data || id != "OcTree" || !(res > 0.0)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "malformed .bt header");
  *res_out = res;
  k3->clear();
  val->clear();
  if (size == 0 || pos >= in.size()) return VMAP_OK;
  struct Frame { int d; uint32_t x, y, z; };
  std::vector<Frame> stack;
  stack.push_back(Frame{0, 0, 0, 0});
  const float cmax = logodds(h->params.threshold_max), cmin = logodds(h->params.threshold_min);
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    if (pos + 2 > in.size()) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "truncated .bt data");
    const uint32_t bits = (uint8_t)in[pos] | ((uint32_t)(uint8_t)in[pos + 1] << 8);
    pos += 2;
    const int cd = f.d + 1;
    const uint32_t bit = 1u << (16 - cd);
    Frame kids[8];
    int n_kids = 0;
    for (uint32_t i = 0; i < 8; i++) {
      const uint32_t two = (bits >> (2 * i)) & 3u;
      if (!two) continue;
      const uint32_t cx = f.x | ((i & 1u) ? bit : 0u), cy = f.y | ((i & 2u) ? bit : 0u), cz = f.z | ((i & 4u) ? bit : 0u);
      if (two == 3u) {
        if (cd >= 16) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "inner node below the finest level in .bt data");
        kids[n_kids++] = Frame{cd, cx, cy, cz};
        continue;
      }
      const float v = (two == 2u) ? cmax : cmin;
      const uint32_t span = 1u << (16 - cd);
      if ((double)span * span * span > 16777216.0 || k3->size() / 3 + (size_t)span * span * span > (1ull << 31))
        return fail(h, VMAP_ERR_CAPACITY, "pruned leaf too coarse to expand into finest-level voxels");
      for (uint32_t z = 0; z < span; z++)
        for (uint32_t y = 0; y < span; y++)
          for (uint32_t x = 0; x < span; x++) {
            k3->push_back((uint16_t)(cx + x)); k3->push_back((uint16_t)(cy + y)); k3->push_back((uint16_t)(cz + z));
            val->push_back(v);
          }
    }
    for (int k = n_kids - 1; k >= 0; k--) stack.push_back(kids[k]);
  }
  return VMAP_OK;
}

int load_bt(vmap* h, const std::string& data) {
  double res = 0.0;
  std::vector<uint16_t> k3;
  std::vector<float> val;
  VM_TRY(parse_bt(h, data, &res, &k3, &val));
  // readBinary replaces the tree, resolution included
  vmap_params p = h->params;
  p.resolution = res;
  VM_TRY(vmap_set_params(h, &p));
  VM_TRY(vmap_reset(h));
  return vmap_import_leaves(h, k3.data(), val.data(), val.size());
}

}  // namespace

extern "C" {

// OcTree::write (full tree, float log-odds per node) into memory: getOctomapFullMsg's payload is
// the stream without the header (with_header = 0), a ".ot" file the stream with it.
int vmap_serialize_ot(vmap* h, int with_header, uint8_t* buffer, size_t capacity, size_t* n_bytes) try {
  if (!h || !n_bytes) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  std::vector<std::pair<uint64_t, float> > leaves;
  VM_TRY(read_leaves_morton(h, &leaves));
  PrunedTree t;
  build_pruned_tree(h, leaves, false, &t);
  std::string out = with_header ? ot_header(h, t.n_nodes) : std::string();
  write_ot_nodes(t, &out);
  *n_bytes = out.size();
  if (!buffer) return VMAP_OK;
  if (capacity < out.size()) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "buffer too small; *n_bytes holds the stream size");
  std::memcpy(buffer, out.data(), out.size());
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// setOctomapFromMsg with a full message (octomap_world.cc:833-844) / OcTree::read: replaces the map.
int vmap_deserialize_ot(vmap* h, int with_header, const uint8_t* buffer, size_t n_bytes) try {
  if (!h || (!buffer && n_bytes)) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  double res = 0.0;
  std::vector<uint16_t> k3;
  std::vector<float> val;
  VM_TRY(parse_ot(h, std::string(reinterpret_cast<const char*>(buffer), n_bytes), with_header != 0, &res, &k3, &val));
  vmap_params p = h->params;
  p.resolution = res;
  VM_TRY(vmap_set_params(h, &p));
  VM_TRY(vmap_reset(h));
  return vmap_import_leaves(h, k3.data(), val.data(), val.size());
} catch (...) { return vmap_on_exception(h); }

// OctomapWorld::setLogOddsBoundingBox (octomap_world.cc:781-818), which is what setFree /
// setOccupied (:497-523) call: every grid position of the adjusted box (adjustBoundingBox,
// :1175-1224) gets its voxel set to log_odds_value (OccupancyOcTreeBase::setNodeValue clamps the
// value to [clamping_thres_min, clamping_thres_max] first; out-of-bounds positions are ignored).  The loops run on the host in the reference's arithmetic (accumulating
// doubles, float narrowing at point3d); the voxels are written on the device in one batch.
int vmap_set_log_odds_bbox(vmap* h, const double* positions_xyz, size_t n_positions,
                           const double bounding_box_size[3], double log_odds_value, int bound_handling) try {
  if (!h || (!positions_xyz && n_positions) || !bounding_box_size) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (bound_handling < 0 || bound_handling > 2) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bound_handling must be 0 (default), 1 (include partial) or 2 (ignore partial)");
  VM_CUDA(h, cudaSetDevice(h->device));
  const double resolution = h->params.resolution;
  const double res_factor = 1.0 / resolution;
  const double epsilon = 0.001;
  auto coord_to_key = [&](float c) -> uint16_t {   // OcTreeBaseImpl::coordToKey (unchecked)
    const double f = std::floor(res_factor * (double)c);
    const int v = (f >= -2147483648.0 && f < 2147483648.0) ? (int)f : (int)0x80000000;
    return (uint16_t)(v + 32768);
  };
  auto key_to_coord = [&](uint16_t k) -> float {   // keyToCoord(OcTreeKey) narrows to float
    return (float)((double((int)k - 32768) + 0.5) * resolution);
  };
  std::vector<uint16_t> k3;
  std::vector<float> val;
  // setNodeValue(point3d, float log_odds_value, lazy): the double is narrowed to float at the call,
  // then clamped: std::min(std::max(v, clamping_thres_min), clamping_thres_max)
  const float value = std::min(std::max((float)log_odds_value, h->P.cmin), h->P.cmax);
  for (size_t p = 0; p < n_positions; p++) {
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; a++) {
      const double pos = positions_xyz[3 * p + a], half = bounding_box_size[a] / 2;
      if (bound_handling == 0) {
        bmin[a] = pos - half - epsilon;
        bmax[a] = pos + half + epsilon;
      } else if (bound_handling == 1) {
        bmin[a] = pos - half + epsilon;
        bmax[a] = pos + half - epsilon;
        bmin[a] = (double)key_to_coord(coord_to_key((float)bmin[a]));
        bmax[a] = (double)key_to_coord(coord_to_key((float)bmax[a]));
        bmin[a] -= epsilon;
        bmax[a] += epsilon;
      } else {
        bmin[a] = pos - half - epsilon;
        bmax[a] = pos + half + epsilon;
        bmin[a] = (double)key_to_coord(coord_to_key((float)bmin[a]));
        bmax[a] = (double)key_to_coord(coord_to_key((float)bmax[a]));
        bmin[a] += resolution;
        bmax[a] -= resolution;
        bmin[a] -= epsilon;
        bmax[a] += epsilon;
      }
    }
    for (double x = bmin[0]; x <= bmax[0]; x += resolution)
      for (double y = bmin[1]; y <= bmax[1]; y += resolution)
        for (double z = bmin[2]; z <= bmax[2]; z += resolution) {
          const float pt[3] = {(float)x, (float)y, (float)z};
          uint16_t key[3];
          bool ok = true;
          for (int a = 0; a < 3; a++) {   // setNodeValue(point3d): coordToKeyChecked or nothing
            const double f = std::floor(res_factor * (double)pt[a]);
            if (!(f >= -32768.0 && f < 32768.0)) { ok = false; break; }
            key[a] = (uint16_t)((int)f + 32768);
          }
          if (!ok) continue;
          k3.push_back(key[0]); k3.push_back(key[1]); k3.push_back(key[2]);
          val.push_back(value);
          if (val.size() > (1u << 28)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bounding box too large for the resolution");
        }
  }
  return vmap_import_leaves(h, k3.data(), val.data(), val.size());
} catch (...) { return vmap_on_exception(h); }

// ---------------------------------------------------------------------------------------------
// The same tree fold / stream code on caller-supplied leaves, without a device: used by hosts that
// already hold the leaves (and by the CPU test-suite, which checks these writers and readers
// against the oracle's pointer-tree implementation on maps the oracle built).
//   kind 0: ".bt" (max-likelihood, OcTree::writeBinaryConst)   1: full-tree data (writeData)
//   kind 2: ".ot" stream with header (OcTree::write)
// ---------------------------------------------------------------------------------------------
int vmap_host_tree_stream(const vmap_params* params, const uint16_t* keys_k3, const float* logodds_in, size_t n,
                          int kind, uint8_t* buffer, size_t capacity, size_t* n_bytes, size_t* n_nodes) try {
  if (!params || ((!keys_k3 || !logodds_in) && n) || !n_bytes || kind < 0 || kind > 2) return VMAP_ERR_INVALID_ARGUMENT;
  std::unique_ptr<vmap> tmp(new vmap());
  tmp->params = *params;
  derive_params(tmp.get());
  std::vector<std::pair<uint64_t, float> > leaves(n);
  for (size_t i = 0; i < n; i++)
    leaves[i] = std::make_pair(host_spread3(keys_k3[3 * i]) | (host_spread3(keys_k3[3 * i + 1]) << 1) | (host_spread3(keys_k3[3 * i + 2]) << 2), logodds_in[i]);
  std::sort(leaves.begin(), leaves.end(), [](const std::pair<uint64_t, float>& a, const std::pair<uint64_t, float>& b) { return a.first < b.first; });
  PrunedTree t;
  build_pruned_tree(tmp.get(), leaves, kind == 0, &t);
  std::string out;
  if (kind == 0) {
    out = bt_header(tmp.get(), t.n_nodes);
    write_bt_nodes(tmp.get(), t, &out);
  } else {
    if (kind == 2) out = ot_header(tmp.get(), t.n_nodes);
    write_ot_nodes(t, &out);
  }
  *n_bytes = out.size();
  if (n_nodes) *n_nodes = t.n_nodes;
  if (!buffer) return VMAP_OK;
  if (capacity < out.size()) return VMAP_ERR_BUFFER_TOO_SMALL;
  std::memcpy(buffer, out.data(), out.size());
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

// Inverse: finest-level leaves of a stream (pruned leaves expanded).  *n_leaves always receives
// the leaf count; the arrays are filled when `capacity` suffices.
int vmap_host_parse_stream(const vmap_params* params, int kind, const uint8_t* buffer, size_t n_bytes,
                           uint16_t* keys_k3, float* logodds_out, size_t capacity, size_t* n_leaves,
                           double* resolution) try {
  if (!params || (!buffer && n_bytes) || !n_leaves || kind < 0 || kind > 2) return VMAP_ERR_INVALID_ARGUMENT;
  std::unique_ptr<vmap> tmp(new vmap());
  tmp->params = *params;
  derive_params(tmp.get());
  std::vector<uint16_t> k3;
  std::vector<float> val;
  double res = params->resolution;
  const std::string in(reinterpret_cast<const char*>(buffer), n_bytes);
  const int st = (kind == 0) ? parse_bt(tmp.get(), in, &res, &k3, &val) : parse_ot(tmp.get(), in, kind == 2, &res, &k3, &val);
  if (st != VMAP_OK) {
    g_create_error = tmp->err;
    return st;
  }
  *n_leaves = val.size();
  if (resolution) *resolution = res;
  if (!keys_k3 || !logodds_out || capacity < val.size()) return val.empty() ? VMAP_OK : VMAP_ERR_BUFFER_TOO_SMALL;
  std::memcpy(keys_k3, k3.data(), k3.size() * sizeof(uint16_t));
  std::memcpy(logodds_out, val.data(), val.size() * sizeof(float));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

// octree_->prune(); octree_->size(): node count of the canonical pruned tree.
int vmap_tree_size(vmap* h, int max_likelihood, size_t* n_nodes) try {
  if (!h || !n_nodes) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  std::vector<std::pair<uint64_t, float> > leaves;
  VM_TRY(read_leaves_morton(h, &leaves));
  PrunedTree t;
  build_pruned_tree(h, leaves, max_likelihood != 0, &t);
  *n_nodes = t.n_nodes;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// OctomapWorld::writeOctomapToBinaryConst (octomap_world.cc:854-856) into memory.
int vmap_serialize_bt(vmap* h, uint8_t* buffer, size_t capacity, size_t* n_bytes) try {
  if (!h || !n_bytes) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  std::string out;
  VM_TRY(serialize_bt(h, &out, nullptr));
  *n_bytes = out.size();
  if (!buffer) return VMAP_OK;   // size query
  if (capacity < out.size()) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "buffer too small; *n_bytes holds the stream size");
  std::memcpy(buffer, out.data(), out.size());
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// OctomapWorld::writeOctomapToFile (octomap_world.cc:849-852) -> OcTree::writeBinary(filename),
// which thresholds the LIVE tree to the clamping values (toMaxLikelihood) and prunes it before
// writing; threshold_live_map != 0 reproduces that side effect on the device map.
int vmap_write_bt(vmap* h, const char* path, int threshold_live_map) try {
  if (!h || !path) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));   // the thresholding pass must not race an apply round still in flight
  if (threshold_live_map) {
    k_to_max_likelihood<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->P, h->m);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
  }
  std::string out;
  VM_TRY(serialize_bt(h, &out, nullptr));
  std::ofstream f(path, std::ios::binary);
  if (!f) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "cannot open output file");
  f.write(out.data(), (std::streamsize)out.size());
  return f.good() ? VMAP_OK : fail(h, VMAP_ERR_INVALID_ARGUMENT, "write failed");
} catch (...) { return vmap_on_exception(h); }

// OctomapWorld::loadOctomapFromFile (octomap_world.cc:846-848) -> readBinary: replaces the map.
int vmap_load_bt(vmap* h, const char* path) try {
  if (!h || !path) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "cannot open input file");
  std::stringstream ss;
  ss << f.rdbuf();
  return load_bt(h, ss.str());
} catch (...) { return vmap_on_exception(h); }

// setOctomapFromMsg with a binary message (octomap_world.cc:833-844): same stream from memory.
int vmap_deserialize_bt(vmap* h, const uint8_t* buffer, size_t n_bytes) try {
  if (!h || (!buffer && n_bytes)) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  return load_bt(h, std::string(reinterpret_cast<const char*>(buffer), n_bytes));
} catch (...) { return vmap_on_exception(h); }

}  // extern "C"
