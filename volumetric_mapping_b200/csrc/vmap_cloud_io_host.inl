// vmap_cloud_io_host.inl -- host side of OctomapManager::loadOctomapCallback's point-cloud branch
// (octomap_manager.cc:313-344): read the xyz of a .pcd / .ply file the way
// pcl::io::loadPCDFile / loadPLYFile<pcl::PointXYZ> would, turn every point whose key is inside
// the map into an occupied key and apply them with an empty free set (updateOccupancy).
// Included by vmap.cu.  Supported: PCD "ascii" and "binary" (x y z as 4- or 8-byte floats among
// any other fields), PLY "ascii", "binary_little_endian" and "binary_big_endian" with scalar
// vertex properties.  PCD "binary_compressed" and PLY list properties inside the vertex element
// are reported as errors, not guessed at.

#include <cctype>
#include <fstream>
#include <sstream>

namespace cloudio {

struct Field { std::string name; int size; char type; int count; };

inline std::string lower(std::string s) {
  for (char& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
inline std::vector<std::string> split(const std::string& line) {
  std::vector<std::string> out;
  std::istringstream is(line);
  std::string t;
  while (is >> t) out.push_back(t);
  return out;
}
inline bool read_line(const std::string& d, size_t* pos, std::string* line) {
  if (*pos >= d.size()) return false;
  size_t e = d.find('\n', *pos);
  if (e == std::string::npos) e = d.size();
  *line = d.substr(*pos, e - *pos);
  if (!line->empty() && line->back() == '\r') line->pop_back();
  *pos = e + 1;
  return true;
}
template <class T> inline T load_as(const char* p, bool swap) {
  unsigned char b[sizeof(T)];
  std::memcpy(b, p, sizeof(T));
  if (swap) std::reverse(b, b + sizeof(T));
  T v;
  std::memcpy(&v, b, sizeof(T));
  return v;
}
// value of a scalar of the given (size, kind: 'F' float, 'I' signed, 'U' unsigned) as float
inline float scalar_as_float(const char* p, int size, char kind, bool swap) {
  if (kind == 'F') return size == 8 ? (float)load_as<double>(p, swap) : load_as<float>(p, swap);
  if (kind == 'I') {
    if (size == 1) return (float)load_as<int8_t>(p, swap);
    if (size == 2) return (float)load_as<int16_t>(p, swap);
    if (size == 4) return (float)load_as<int32_t>(p, swap);
    return (float)load_as<int64_t>(p, swap);
  }
  if (size == 1) return (float)load_as<uint8_t>(p, swap);
  if (size == 2) return (float)load_as<uint16_t>(p, swap);
  if (size == 4) return (float)load_as<uint32_t>(p, swap);
  return (float)load_as<uint64_t>(p, swap);
}
inline bool host_is_big_endian() {
  const uint16_t one = 1;
  return *reinterpret_cast<const unsigned char*>(&one) == 0;
}

inline int read_pcd(const std::string& d, std::vector<float>* xyz, std::string* err) {
  size_t pos = 0;
  std::string line;
  std::vector<std::string> names, types;
  std::vector<int> sizes, counts;
  long long points = -1, width = -1, height = 1;
  std::string data_kind;
  while (read_line(d, &pos, &line)) {
    if (line.empty() || line[0] == '#') continue;
    const std::vector<std::string> t = split(line);
    if (t.empty()) continue;
    const std::string key = lower(t[0]);
    if (key == "fields" || key == "columns") names.assign(t.begin() + 1, t.end());
    else if (key == "size") for (size_t i = 1; i < t.size(); i++) sizes.push_back(std::atoi(t[i].c_str()));
    else if (key == "type") types.assign(t.begin() + 1, t.end());
    else if (key == "count") for (size_t i = 1; i < t.size(); i++) counts.push_back(std::atoi(t[i].c_str()));
    else if (key == "width" && t.size() > 1) width = std::atoll(t[1].c_str());
    else if (key == "height" && t.size() > 1) height = std::atoll(t[1].c_str());
    else if (key == "points" && t.size() > 1) points = std::atoll(t[1].c_str());
    else if (key == "data" && t.size() > 1) { data_kind = lower(t[1]); break; }
  }
  if (data_kind.empty()) { *err = "PCD: no DATA line"; return VMAP_ERR_INVALID_ARGUMENT; }
  if (names.empty()) { *err = "PCD: no FIELDS line"; return VMAP_ERR_INVALID_ARGUMENT; }
  if (sizes.empty()) sizes.assign(names.size(), 4);
  if (types.empty()) types.assign(names.size(), "F");
  if (counts.empty()) counts.assign(names.size(), 1);
  if (sizes.size() != names.size() || types.size() != names.size() || counts.size() != names.size()) {
    *err = "PCD: FIELDS / SIZE / TYPE / COUNT disagree";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  if (points < 0) points = width > 0 ? width * height : -1;
  if (points < 0) { *err = "PCD: neither POINTS nor WIDTH given"; return VMAP_ERR_INVALID_ARGUMENT; }
  int idx[3] = {-1, -1, -1}, offset[3] = {0, 0, 0}, column[3] = {0, 0, 0};
  int step = 0, col = 0;
  for (size_t f = 0; f < names.size(); f++) {
    const std::string n = lower(names[f]);
    for (int a = 0; a < 3; a++)
      if (n == std::string(1, (char)('x' + a))) { idx[a] = (int)f; offset[a] = step; column[a] = col; }
    step += sizes[f] * counts[f];
    col += counts[f];
  }
  if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0) { *err = "PCD: x / y / z fields missing"; return VMAP_ERR_INVALID_ARGUMENT; }
  for (int a = 0; a < 3; a++) {
    const char ty = (char)std::toupper((unsigned char)types[idx[a]][0]);
    if (ty != 'F' || (sizes[idx[a]] != 4 && sizes[idx[a]] != 8)) { *err = "PCD: x / y / z must be float fields"; return VMAP_ERR_INVALID_ARGUMENT; }
  }
  xyz->resize((size_t)points * 3);
  if (data_kind == "ascii") {
    for (long long i = 0; i < points; i++) {
      do {
        if (!read_line(d, &pos, &line)) { *err = "PCD: fewer points than POINTS says"; return VMAP_ERR_INVALID_ARGUMENT; }
      } while (line.empty());
      const std::vector<std::string> t = split(line);
      if ((int)t.size() < col) { *err = "PCD: short data line"; return VMAP_ERR_INVALID_ARGUMENT; }
      for (int a = 0; a < 3; a++) (*xyz)[3 * i + a] = (float)std::strtod(t[column[a]].c_str(), nullptr);   // strtod reads "nan" / "inf"
    }
  } else if (data_kind == "binary") {
    if (d.size() - std::min(pos, d.size()) < (size_t)points * (size_t)step) { *err = "PCD: truncated binary data"; return VMAP_ERR_INVALID_ARGUMENT; }
    const bool swap = host_is_big_endian();
    for (long long i = 0; i < points; i++)
      for (int a = 0; a < 3; a++)
        (*xyz)[3 * i + a] = scalar_as_float(d.data() + pos + (size_t)i * step + offset[a], sizes[idx[a]], 'F', swap);
  } else {
    *err = "PCD: DATA " + data_kind + " is not supported (ascii and binary are)";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  return VMAP_OK;
}

inline bool ply_scalar(const std::string& t, int* size, char* kind) {
  static const struct { const char* n; int s; char k; } tab[] = {
      {"char", 1, 'I'}, {"int8", 1, 'I'}, {"uchar", 1, 'U'}, {"uint8", 1, 'U'}, {"short", 2, 'I'}, {"int16", 2, 'I'},
      {"ushort", 2, 'U'}, {"uint16", 2, 'U'}, {"int", 4, 'I'}, {"int32", 4, 'I'}, {"uint", 4, 'U'}, {"uint32", 4, 'U'},
      {"float", 4, 'F'}, {"float32", 4, 'F'}, {"double", 8, 'F'}, {"float64", 8, 'F'}};
  for (const auto& e : tab)
    if (t == e.n) { *size = e.s; *kind = e.k; return true; }
  return false;
}

inline int read_ply(const std::string& d, std::vector<float>* xyz, std::string* err) {
  size_t pos = 0;
  std::string line;
  if (!read_line(d, &pos, &line) || lower(line).substr(0, 3) != "ply") { *err = "PLY: missing magic"; return VMAP_ERR_INVALID_ARGUMENT; }
  struct Prop { std::string name; int size; char kind; bool list; };
  struct Elem { std::string name; long long count; std::vector<Prop> props; };
  std::vector<Elem> elems;
  std::string format;
  bool ended = false;
  while (read_line(d, &pos, &line)) {
    const std::vector<std::string> t = split(line);
    if (t.empty()) continue;
    const std::string key = lower(t[0]);
    if (key == "format" && t.size() > 1) format = lower(t[1]);
    else if (key == "element" && t.size() > 2) elems.push_back(Elem{lower(t[1]), std::atoll(t[2].c_str()), {}});
    else if (key == "property" && !elems.empty() && t.size() > 2) {
      Prop p{lower(t.back()), 0, 'F', false};
      if (lower(t[1]) == "list") p.list = true;
      else if (!ply_scalar(lower(t[1]), &p.size, &p.kind)) { *err = "PLY: unknown property type " + t[1]; return VMAP_ERR_INVALID_ARGUMENT; }
      elems.back().props.push_back(p);
    } else if (key == "end_header") { ended = true; break; }
  }
  if (!ended) { *err = "PLY: no end_header"; return VMAP_ERR_INVALID_ARGUMENT; }
  const bool ascii = format == "ascii";
  if (!ascii && format != "binary_little_endian" && format != "binary_big_endian") { *err = "PLY: unknown format " + format; return VMAP_ERR_INVALID_ARGUMENT; }
  const bool swap = !ascii && ((format == "binary_big_endian") != host_is_big_endian());
  for (size_t e = 0; e < elems.size(); e++) {
    const Elem& el = elems[e];
    const bool vertex = el.name == "vertex";
    int at[3] = {-1, -1, -1};
    for (size_t p = 0; p < el.props.size(); p++) {
      if (el.props[p].list) { *err = "PLY: list property in or before the vertex element"; return VMAP_ERR_INVALID_ARGUMENT; }
      for (int a = 0; a < 3; a++)
        if (el.props[p].name == std::string(1, (char)('x' + a))) at[a] = (int)p;
    }
    if (vertex && (at[0] < 0 || at[1] < 0 || at[2] < 0)) { *err = "PLY: vertex element without x / y / z"; return VMAP_ERR_INVALID_ARGUMENT; }
    if (vertex) xyz->resize((size_t)el.count * 3);
    for (long long i = 0; i < el.count; i++) {
      if (ascii) {
        do {
          if (!read_line(d, &pos, &line)) { *err = "PLY: fewer elements than the header says"; return VMAP_ERR_INVALID_ARGUMENT; }
        } while (split(line).empty());
        if (!vertex) continue;
        const std::vector<std::string> t = split(line);
        if (t.size() < el.props.size()) { *err = "PLY: short vertex line"; return VMAP_ERR_INVALID_ARGUMENT; }
        for (int a = 0; a < 3; a++) (*xyz)[3 * i + a] = (float)std::strtod(t[at[a]].c_str(), nullptr);
      } else {
        size_t off = 0;
        for (size_t p = 0; p < el.props.size(); p++) {
          if (pos + off + (size_t)el.props[p].size > d.size()) { *err = "PLY: truncated binary data"; return VMAP_ERR_INVALID_ARGUMENT; }
          if (vertex)
            for (int a = 0; a < 3; a++)
              if (at[a] == (int)p) (*xyz)[3 * i + a] = scalar_as_float(d.data() + pos + off, el.props[p].size, el.props[p].kind, swap);
          off += (size_t)el.props[p].size;
        }
        pos += off;
      }
    }
    if (vertex) return VMAP_OK;
  }
  *err = "PLY: no vertex element";
  return VMAP_ERR_INVALID_ARGUMENT;
}

inline int read_cloud_file(const char* path, std::vector<float>* xyz, std::string* err) {
  const std::string p(path);
  const size_t dot = p.find_last_of('.');
  const std::string ext = dot == std::string::npos ? "" : p.substr(dot + 1);   // (case-sensitive like the reference)
  if (ext != "pcd" && ext != "ply") { *err = "No known file extension (.pcd, .ply): " + p; return VMAP_ERR_INVALID_ARGUMENT; }
  std::ifstream f(path, std::ios::binary);
  if (!f) { *err = "cannot open " + p; return VMAP_ERR_INVALID_ARGUMENT; }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string d = ss.str();
  return ext == "pcd" ? read_pcd(d, xyz, err) : read_ply(d, xyz, err);
}

}  // namespace cloudio

extern "C" {

// xyz of a .pcd / .ply file (pure host; *n always receives the point count).
int vmap_host_read_pointcloud_file(const char* path, float* xyz, size_t capacity, size_t* n) try {
  if (!path || !n) return VMAP_ERR_INVALID_ARGUMENT;
  std::vector<float> pts;
  std::string err;
  const int st = cloudio::read_cloud_file(path, &pts, &err);
  if (st != VMAP_OK) {
    g_create_error = err;
    return st;
  }
  *n = pts.size() / 3;
  if (!xyz || capacity < *n) return *n ? VMAP_ERR_BUFFER_TOO_SMALL : VMAP_OK;
  std::memcpy(xyz, pts.data(), pts.size() * sizeof(float));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

// The occupied-only insertion of loadOctomapCallback (octomap_manager.cc:330-341): every point
// with coordToKeyChecked(point) inside the map becomes an occupied key (a KeySet: duplicates
// collapse), the free set stays empty, updateOccupancy applies them.
int vmap_insert_occupied_points(vmap* h, const float* xyz, size_t n, size_t stride_bytes) try {
  if (!h || (!xyz && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  std::vector<uint16_t> occ;
  occ.reserve(3 * n);
  const double rf = 1.0 / h->params.resolution;
  for (size_t i = 0; i < n; i++) {
    const float* p = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xyz) + i * stride_bytes);
    int s[3];
    bool ok = true;
    for (int a = 0; a < 3; a++) {
      // coordToKeyChecked: (int)floor(res_factor * x) + 32768, x86 conversion (NaN / huge -> INT_MIN)
      const double f = std::floor(rf * (double)p[a]);
      const int q = (f >= -2147483648.0 && f < 2147483648.0) ? (int)f : (int)0x80000000;
      s[a] = (int)((unsigned)q + 32768u);
      ok = ok && s[a] >= 0 && s[a] < 65536;
    }
    if (!ok) continue;
    occ.push_back((uint16_t)s[0]); occ.push_back((uint16_t)s[1]); occ.push_back((uint16_t)s[2]);
  }
  return vmap_apply_keys(h, nullptr, 0, occ.data(), occ.size() / 3);
} catch (...) { return vmap_on_exception(h); }

// loadOctomapCallback's non-.bt branch.
int vmap_load_occupied_pointcloud_file(vmap* h, const char* path, size_t* n_points) try {
  if (!h || !path) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  std::vector<float> pts;
  std::string err;
  const int st = cloudio::read_cloud_file(path, &pts, &err);
  if (st != VMAP_OK) return fail(h, st, err.c_str());
  if (n_points) *n_points = pts.size() / 3;
  return vmap_insert_occupied_points(h, pts.data(), pts.size() / 3, 12);
} catch (...) { return vmap_on_exception(h); }

}  // extern "C"
