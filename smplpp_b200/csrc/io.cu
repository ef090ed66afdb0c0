// Host-side data formats on either side of the hot path; no CUDA in this file except the
// model / decoder creation calls the loaders end with.
//
//   * JSON parameter files: the SMPL model (keys of src/SMPL.cpp:572-612, written by scripts/preprocess.py:109-117) and
//     the VPoser decoder (keys of src/VPoser.cpp:185-237).  The reference parses them with nlohmann::json +
//     xt::from_json (both un-vendored); a parameter file is one flat object of nested numeric arrays, so a 100-line
//     recursive-descent reader that keeps only numbers and shapes replaces both.
//   * C3D motion capture files: what the node reads through ezc3d (un-vendored; node/node.cpp:572-595, 667-691):
//     POINT:LABELS (suffix match against the task names), header().frameRate() / nbFrames(), and per frame and point
//     x, y, z, isEmpty().  Layout per the public C3D specification (c3d.org): 512-byte header block, parameter section,
//     frame-major point data as 4 x float32 (scale < 0) or 4 x int16 (scaled by POINT:SCALE), Intel byte order.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"

using namespace sb;

// ------------------------------------------------------------------------------------------------------------
// JSON: one object of (nested) numeric arrays
// ------------------------------------------------------------------------------------------------------------
struct smplpp_json
{
  struct Array
  {
    std::vector<int64_t> shape;
    std::vector<double> data;
    bool ragged = false;
  };
  std::map<std::string, Array> arrays;
};

namespace
{
struct JsonReader
{
  const char * p;
  const char * end;
  std::string err;

  void ws()
  {
    while(p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
  }
  bool fail_at(const char * what)
  {
    if(err.empty()) err = what;
    return false;
  }
  bool string(std::string & out)
  {
    if(p >= end || *p != '"') return fail_at("expected a string");
    p++;
    out.clear();
    while(p < end && *p != '"')
    {
      if(*p == '\\' && p + 1 < end) p++;
      out.push_back(*p++);
    }
    if(p >= end) return fail_at("unterminated string");
    p++;
    return true;
  }
  // numeric array of any depth -> shape + flat row-major data; depth = nesting level of this '['
  bool array(smplpp_json::Array & a, size_t depth)
  {
    p++; // '['
    int64_t count = 0;
    ws();
    if(p < end && *p == ']')
    {
      p++;
    }
    else
    {
      for(;;)
      {
        ws();
        if(p >= end) return fail_at("unterminated array");
        if(*p == '[')
        {
          if(!array(a, depth + 1)) return false;
        }
        else
        {
          char * stop = nullptr;
          const double v = strtod(p, &stop);
          if(stop == p) return fail_at("expected a number");
          p = stop;
          a.data.push_back(v);
          if(a.shape.size() < depth + 1) a.shape.resize(depth + 1, -1);
        }
        count++;
        ws();
        if(p < end && *p == ',')
        {
          p++;
          continue;
        }
        if(p < end && *p == ']')
        {
          p++;
          break;
        }
        return fail_at("expected ',' or ']'");
      }
    }
    if(a.shape.size() < depth + 1) a.shape.resize(depth + 1, -1);
    if(a.shape[depth] == -1)
      a.shape[depth] = count;
    else if(a.shape[depth] != count)
      a.ragged = true;
    return true;
  }
  // any value we do not keep (strings, literals, nested objects)
  bool skip_value()
  {
    ws();
    if(p >= end) return fail_at("unexpected end");
    if(*p == '"')
    {
      std::string s;
      return string(s);
    }
    if(*p == '{')
    {
      p++;
      ws();
      if(p < end && *p == '}')
      {
        p++;
        return true;
      }
      for(;;)
      {
        ws();
        std::string k;
        if(!string(k)) return false;
        ws();
        if(p >= end || *p != ':') return fail_at("expected ':'");
        p++;
        if(!skip_value()) return false;
        ws();
        if(p < end && *p == ',')
        {
          p++;
          continue;
        }
        if(p < end && *p == '}')
        {
          p++;
          return true;
        }
        return fail_at("expected ',' or '}'");
      }
    }
    if(*p == '[')
    {
      smplpp_json::Array tmp;
      return array(tmp, 0);
    }
    while(p < end && *p != ',' && *p != '}' && *p != ']') p++; // number / true / false / null
    return true;
  }
  bool object(smplpp_json & out)
  {
    ws();
    if(p >= end || *p != '{') return fail_at("a parameter file is one JSON object");
    p++;
    ws();
    if(p < end && *p == '}') return true;
    for(;;)
    {
      ws();
      std::string key;
      if(!string(key)) return false;
      ws();
      if(p >= end || *p != ':') return fail_at("expected ':'");
      p++;
      ws();
      if(p < end && *p == '[')
      {
        smplpp_json::Array a;
        if(!array(a, 0)) return false;
        out.arrays[key] = std::move(a);
      }
      else if(p < end && (*p == '-' || (*p >= '0' && *p <= '9')))
      {
        char * stop = nullptr;
        smplpp_json::Array a;
        a.data.push_back(strtod(p, &stop)); // scalar: shape ()
        p = stop;
        out.arrays[key] = std::move(a);
      }
      else if(!skip_value())
        return false;
      ws();
      if(p < end && *p == ',')
      {
        p++;
        continue;
      }
      if(p < end && *p == '}') return true;
      return fail_at("expected ',' or '}'");
    }
  }
};

bool read_file(const char * path, std::string & out)
{
  std::ifstream f(path, std::ios::binary);
  if(!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

int load_json(const char * path, const char * module, const char * missing_msg, std::unique_ptr<smplpp_json> & out)
{
  std::string text;
  if(!path || !read_file(path, text)) return fail(SMPLPP_ERR_IO, module, missing_msg);
  out.reset(new smplpp_json());
  JsonReader r{text.data(), text.data() + text.size(), {}};
  if(!r.object(*out)) return fail(SMPLPP_ERR_IO, module, std::string("Cannot parse the JSON file: ") + r.err);
  return SMPLPP_OK;
}

template<typename T>
std::vector<T> cast_to(const std::vector<double> & v)
{
  std::vector<T> o(v.size());
  for(size_t i = 0; i < v.size(); i++) o[i] = static_cast<T>(v[i]);
  return o;
}

const smplpp_json::Array * find(const smplpp_json & j, const char * key)
{
  auto it = j.arrays.find(key);
  if(it == j.arrays.end() || it->second.ragged) return nullptr;
  // rectangular means data.size() == prod(shape): a mixed-depth array such as [[1,2],3] has consistent per-level counts
  // (2 and 2) but only 3 values, and every consumer indexes data by the shape
  const smplpp_json::Array & a = it->second;
  size_t count = 1;
  for(int64_t d : a.shape)
  {
    if(d < 0) return nullptr;
    if(d != 0 && count > a.data.size() / static_cast<size_t>(d)) return nullptr;
    count *= static_cast<size_t>(d);
  }
  return count == a.data.size() ? &a : nullptr;
}
} // namespace

extern "C" int smplpp_json_open(const char * path, smplpp_json_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "JSON", "null output");
  std::unique_ptr<smplpp_json> j;
  const int rc = load_json(path, "JSON", "Cannot find a JSON file!", j);
  if(rc != SMPLPP_OK) return rc;
  *out = j.release();
  return SMPLPP_OK;
}

extern "C" void smplpp_json_close(smplpp_json_t * j)
{
  delete j;
}

extern "C" int smplpp_json_array(const smplpp_json_t * j, const char * key, int32_t * ndim, int64_t * shape8,
                                 const double ** data)
{
  if(!j || !key || !ndim || !shape8 || !data) return fail(SMPLPP_ERR_INVALID, "JSON", "null argument");
  const smplpp_json::Array * a = find(*j, key);
  if(!a) return fail(SMPLPP_ERR_IO, "JSON", std::string("no rectangular numeric array under key ") + key);
  if(a->shape.size() > 8) return fail(SMPLPP_ERR_IO, "JSON", "more than 8 dimensions");
  *ndim = static_cast<int32_t>(a->shape.size());
  for(size_t i = 0; i < a->shape.size(); i++) shape8[i] = a->shape[i];
  *data = a->data.data();
  return SMPLPP_OK;
}

// SMPL::init (src/SMPL.cpp:560-617): same keys, same shape checks, same messages
static int model_from_arrays(const smplpp_json * j, smplpp_model_t ** out);

extern "C" int smplpp_model_load_json(const char * path, smplpp_model_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "SMPL", "Cannot initialize a SMPL model!");
  std::unique_ptr<smplpp_json> j;
  int rc = load_json(path, "SMPL", "Cannot initialize a SMPL model!", j); // SMPL.cpp:614-617: the file does not exist
  if(rc != SMPLPP_OK) return rc;
  return model_from_arrays(j.get(), out);
}

// the .npz twin of the model JSON (scripts/preprocess.py:98-117)
extern "C" int smplpp_model_load_npz(const char * path, smplpp_model_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "SMPL", "Cannot initialize a SMPL model!");
  smplpp_json_t * j = nullptr;
  int rc = smplpp_npz_open(path, &j);
  if(rc != SMPLPP_OK) return rc;
  std::unique_ptr<smplpp_json> hold(j);
  return model_from_arrays(j, out);
}

static int model_from_arrays(const smplpp_json * j, smplpp_model_t ** out)
{
  const char * keys[] = {"face_indices", "shape_blend_shapes", "pose_blend_shapes", "vertices_template", "joint_regressor",
                         "kinematic_tree", "weights"};
  const smplpp_json::Array * a[7];
  for(int i = 0; i < 7; i++)
  {
    a[i] = find(*j, keys[i]);
    if(!a[i]) return fail(SMPLPP_ERR_IO, "SMPL", std::string("Cannot initialize a SMPL model! (key ") + keys[i] + " is missing)");
  }
  const smplpp_json::Array &faces = *a[0], &sbs = *a[1], &pbs = *a[2], &templ = *a[3], &jreg = *a[4], &tree = *a[5], &w = *a[6];
  if(sbs.shape.size() != 3 || sbs.shape[2] != kShapeDim)
    return fail(SMPLPP_ERR_IO, "SMPL",
                "Shape parameter dimensions are invalid: " + std::to_string(sbs.shape.size() == 3 ? sbs.shape[2] : -1) + " != "
                    + std::to_string(kShapeDim)); // SMPL.cpp:579-583
  if(pbs.shape.size() != 3 || pbs.shape[2] != kPoseDim)
    return fail(SMPLPP_ERR_IO, "SMPL",
                "Pose parameter dimensions are invalid: " + std::to_string(pbs.shape.size() == 3 ? pbs.shape[2] : -1) + " != "
                    + std::to_string(kPoseDim)); // SMPL.cpp:586-590
  const int64_t V = templ.shape.empty() ? 0 : templ.shape[0];
  const bool ok = V >= 1 && templ.shape.size() == 2 && templ.shape[1] == 3 && sbs.shape[0] == V && sbs.shape[1] == 3
                  && pbs.shape[0] == V && pbs.shape[1] == 3 && jreg.shape.size() == 2 && jreg.shape[0] == kJoints
                  && jreg.shape[1] == V && tree.shape.size() == 2 && tree.shape[0] == 2 && tree.shape[1] == kJoints
                  && w.shape.size() == 2 && w.shape[0] == V && w.shape[1] == kJoints && faces.shape.size() == 2
                  && faces.shape[1] == 3;
  if(!ok) return fail(SMPLPP_ERR_IO, "SMPL", "Cannot initialize a SMPL model! (array shapes do not match)");
  const std::vector<int32_t> h_faces = cast_to<int32_t>(faces.data);
  const std::vector<float> h_sbs = cast_to<float>(sbs.data), h_pbs = cast_to<float>(pbs.data), h_templ = cast_to<float>(templ.data),
                           h_jreg = cast_to<float>(jreg.data), h_w = cast_to<float>(w.data);
  const std::vector<int64_t> h_tree = cast_to<int64_t>(tree.data); // the root's parent 4294967295 survives the double round trip
  smplpp_model_desc d;
  d.vertex_num = V;
  d.face_num = faces.shape[0];
  d.face_indices = h_faces.data();
  d.shape_blend_shapes = h_sbs.data();
  d.pose_blend_shapes = h_pbs.data();
  d.vertices_template = h_templ.data();
  d.joint_regressor = h_jreg.data();
  d.kinematic_tree = h_tree.data();
  d.weights = h_w.data();
  return smplpp_model_create(&d, out);
}

// VPoserDecoderImpl::loadParamsFromJson (src/VPoser.cpp:169-238): same keys, same shape checks, same messages
extern "C" int smplpp_vposer_load_json(const char * path, smplpp_vposer_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "VPoser", "Cannot find a JSON file!");
  std::unique_ptr<smplpp_json> j;
  int rc = load_json(path, "VPoser", "Cannot find a JSON file!", j); // VPoser.cpp:181-184
  if(rc != SMPLPP_OK) return rc;
  struct Want
  {
    const char * key;
    int64_t d0, d1; // d1 = 0: one-dimensional
  };
  const int64_t hidden = 512, latent = 32, outd = 126;
  const Want want[6] = {{"decoder_net.0.weight", hidden, latent}, {"decoder_net.0.bias", hidden, 0},
                        {"decoder_net.3.weight", hidden, hidden}, {"decoder_net.3.bias", hidden, 0},
                        {"decoder_net.5.weight", outd, hidden},   {"decoder_net.5.bias", outd, 0}};
  std::vector<float> host[6];
  for(int i = 0; i < 6; i++)
  {
    const smplpp_json::Array * a = find(*j, want[i].key);
    const bool ok = a
                    && (want[i].d1 ? (a->shape.size() == 2 && a->shape[0] == want[i].d0 && a->shape[1] == want[i].d1)
                                   : (a->shape.size() == 1 && a->shape[0] == want[i].d0));
    if(!ok) return fail(SMPLPP_ERR_IO, "VPoser", std::string("invalid dimension of ") + want[i].key + " from JSON file!");
    host[i] = cast_to<float>(a->data);
  }
  smplpp_vposer_desc d;
  d.w0 = host[0].data(), d.b0 = host[1].data(), d.w3 = host[2].data(), d.b3 = host[3].data(), d.w5 = host[4].data(),
  d.b5 = host[5].data();
  return smplpp_vposer_create(&d, out);
}

// ------------------------------------------------------------------------------------------------------------
// C3D
// ------------------------------------------------------------------------------------------------------------
struct smplpp_c3d
{
  std::string bytes;      // the whole file
  int points = 0;         // 3D points per frame
  int analog_per_frame = 0; // analog words per frame (samples x channels)
  int64_t frames = 0;
  int first_frame = 1;
  float scale = 0.f;      // < 0: float32 data, else int16 * scale
  float rate = 0.f;
  size_t data_offset = 0;
  std::vector<std::string> labels;       // POINT:LABELS, LABELS2, ... padded to `points` (smplpp_c3d_label)
  std::vector<std::string> labels_param; // the POINT:LABELS values exactly as stored (smplpp_c3d_find_label)
  std::string units;
};

namespace
{
template<typename T>
T rd(const std::string & b, size_t off)
{
  T v;
  memcpy(&v, b.data() + off, sizeof(T));
  return v;
}

struct C3dParam
{
  int type = 0; // -1 char, 1 byte, 2 int16, 4 float
  std::vector<int> dims;
  size_t data_off = 0;
};

// walks the parameter section and returns the parameters of groups POINT and TRIAL, keyed "GROUP:NAME"
bool c3d_parameters(const std::string & b, size_t start, std::map<std::string, C3dParam> & out, std::string & err)
{
  if(start + 4 > b.size()) return err = "parameter section is outside the file", false;
  const int proc = static_cast<uint8_t>(b[start + 3]);
  if(proc != 84) return err = "only Intel (little-endian IEEE) C3D files are supported, processor type " + std::to_string(proc), false;
  std::map<int, std::string> groups;
  struct Pending
  {
    int gid;
    std::string name;
    C3dParam prm;
  };
  std::vector<Pending> params;
  size_t pos = start + 4;
  while(pos + 2 <= b.size())
  {
    const int nlen = std::abs(static_cast<int>(static_cast<int8_t>(b[pos])));
    const int gid = static_cast<int8_t>(b[pos + 1]);
    if(nlen == 0) break;
    if(pos + 2 + nlen + 2 > b.size()) return err = "truncated parameter record", false;
    std::string name(b.data() + pos + 2, nlen);
    for(auto & c : name) c = static_cast<char>(toupper(static_cast<unsigned char>(c)));
    const size_t off_field = pos + 2 + nlen;
    const int next = rd<int16_t>(b, off_field);
    if(gid < 0)
      groups[-gid] = name;
    else
    {
      size_t q = off_field + 2;
      if(q + 2 > b.size()) return err = "truncated parameter record", false;
      C3dParam prm;
      prm.type = static_cast<int8_t>(b[q]);
      const int nd = static_cast<uint8_t>(b[q + 1]);
      q += 2;
      if(q + nd > b.size()) return err = "truncated parameter record", false;
      for(int i = 0; i < nd; i++) prm.dims.push_back(static_cast<uint8_t>(b[q + i]));
      prm.data_off = q + nd;
      params.push_back({gid, name, prm});
    }
    if(next <= 0) break;
    pos = off_field + next;
  }
  for(auto & pp : params)
  {
    auto g = groups.find(pp.gid);
    if(g != groups.end()) out[g->second + ":" + pp.name] = pp.prm;
  }
  return true;
}

double c3d_scalar(const std::string & b, const C3dParam & prm)
{
  // a truncated parameter section must not be read past the end of the file
  const size_t need = prm.type == 4 ? 4 : (prm.type == 2 ? 2 : (prm.type == 1 ? 1 : 0));
  if(need == 0 || prm.data_off > b.size() || b.size() - prm.data_off < need) return 0.0;
  if(prm.type == 4) return rd<float>(b, prm.data_off);
  if(prm.type == 2) return static_cast<uint16_t>(rd<int16_t>(b, prm.data_off)); // counts are stored unsigned
  return static_cast<uint8_t>(b[prm.data_off]);
}

void c3d_strings(const std::string & b, const C3dParam & prm, std::vector<std::string> & out)
{
  if(prm.type != -1 || prm.dims.empty()) return;
  const int len = prm.dims[0];
  int count = 1;
  for(size_t i = 1; i < prm.dims.size(); i++) count *= prm.dims[i];
  for(int i = 0; i < count; i++)
  {
    if(prm.data_off + static_cast<size_t>(i + 1) * len > b.size()) break;
    std::string s(b.data() + prm.data_off + static_cast<size_t>(i) * len, len);
    while(!s.empty() && (s.back() == ' ' || s.back() == '\0')) s.pop_back();
    out.push_back(s);
  }
}
} // namespace

extern "C" int smplpp_c3d_open(const char * path, smplpp_c3d_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "C3D", "null output");
  std::unique_ptr<smplpp_c3d> c(new smplpp_c3d());
  if(!path || !read_file(path, c->bytes)) return fail(SMPLPP_ERR_IO, "C3D", std::string("Cannot open the C3D file ") + (path ? path : ""));
  const std::string & b = c->bytes;
  if(b.size() < 512 || static_cast<uint8_t>(b[1]) != 0x50) return fail(SMPLPP_ERR_IO, "C3D", "not a C3D file (key byte 0x50 missing)");
  const int param_block = static_cast<uint8_t>(b[0]);
  c->points = rd<uint16_t>(b, 2);
  c->analog_per_frame = rd<uint16_t>(b, 4);
  const int first = rd<uint16_t>(b, 6), last = rd<uint16_t>(b, 8);
  c->first_frame = first;
  c->scale = rd<float>(b, 12);
  int data_block = rd<uint16_t>(b, 16);
  c->rate = rd<float>(b, 20);
  c->frames = static_cast<int64_t>(last) - first + 1;
  std::map<std::string, C3dParam> prm;
  std::string err;
  if(!c3d_parameters(b, static_cast<size_t>(param_block - 1) * 512, prm, err)) return fail(SMPLPP_ERR_IO, "C3D", err);
  // the parameter section is authoritative where the 16-bit header fields overflow (long captures)
  auto it = prm.find("POINT:FRAMES");
  if(it != prm.end())
  {
    const double n = c3d_scalar(b, it->second);
    if(n > 0 && it->second.type == 4) c->frames = static_cast<int64_t>(n);
    else if(n > 0 && c->frames <= 0) c->frames = static_cast<int64_t>(n);
  }
  if((it = prm.find("POINT:DATA_START")) != prm.end() && data_block == 0) data_block = static_cast<int>(c3d_scalar(b, it->second));
  if((it = prm.find("POINT:RATE")) != prm.end() && !(c->rate > 0.f)) c->rate = static_cast<float>(c3d_scalar(b, it->second));
  if((it = prm.find("POINT:SCALE")) != prm.end() && c->scale == 0.f) c->scale = static_cast<float>(c3d_scalar(b, it->second));
  if((it = prm.find("POINT:LABELS")) != prm.end()) c3d_strings(b, it->second, c->labels);
  c->labels_param = c->labels; // what parameters().group("POINT").parameter("LABELS").valuesAsString() returns (node.cpp:582-583)
  for(int k = 2; k < 10; k++) // POINT:LABELS2 ... for more than 255 points
    if((it = prm.find("POINT:LABELS" + std::to_string(k))) != prm.end()) c3d_strings(b, it->second, c->labels);
  if((it = prm.find("POINT:UNITS")) != prm.end())
  {
    std::vector<std::string> u;
    c3d_strings(b, it->second, u);
    if(!u.empty()) c->units = u[0];
  }
  if(c->labels.size() < static_cast<size_t>(c->points)) c->labels.resize(static_cast<size_t>(c->points)); // unnamed points keep an empty label; never shrink
  if(data_block < 1 || c->points < 0 || c->frames < 0) return fail(SMPLPP_ERR_IO, "C3D", "inconsistent C3D header");
  c->data_offset = static_cast<size_t>(data_block - 1) * 512;
  const size_t word = c->scale < 0.f ? 4 : 2;
  const size_t frame_bytes = (static_cast<size_t>(c->points) * 4 + c->analog_per_frame) * word;
  if(c->data_offset + static_cast<size_t>(c->frames) * frame_bytes > b.size())
    return fail(SMPLPP_ERR_IO, "C3D", "the C3D file is shorter than its header says");
  *out = c.release();
  return SMPLPP_OK;
}

extern "C" void smplpp_c3d_close(smplpp_c3d_t * c)
{
  delete c;
}
extern "C" int64_t smplpp_c3d_frame_count(const smplpp_c3d_t * c)
{
  return c ? c->frames : 0;
}
extern "C" int64_t smplpp_c3d_point_count(const smplpp_c3d_t * c)
{
  return c ? c->points : 0;
}
extern "C" double smplpp_c3d_frame_rate(const smplpp_c3d_t * c)
{
  return c ? c->rate : 0.0;
}
extern "C" const char * smplpp_c3d_label(const smplpp_c3d_t * c, int64_t i)
{
  return (c && i >= 0 && i < static_cast<int64_t>(c->labels.size())) ? c->labels[static_cast<size_t>(i)].c_str() : "";
}
extern "C" const char * smplpp_c3d_units(const smplpp_c3d_t * c)
{
  return c ? c->units.c_str() : "";
}

// node/node.cpp:580-594: index of the first label that ENDS with `name` (std::find_if + std::distance: the point
// count when there is none)
extern "C" int64_t smplpp_c3d_find_label(const smplpp_c3d_t * c, const char * name)
{
  if(!c || !name) return 0;
  // std::find_if over the FULL POINT:LABELS value list, std::distance to end() when nothing matches (node.cpp:584-593)
  const size_t n = strlen(name);
  for(size_t i = 0; i < c->labels_param.size(); i++)
  {
    const std::string & s = c->labels_param[i];
    if(s.size() >= n && s.compare(s.size() - n, n, name) == 0) return static_cast<int64_t>(i);
  }
  return static_cast<int64_t>(c->labels_param.size());
}

// frames [first, first + count) -> xyz (count, points, 3) and valid (count, points): 1 when the point exists
// (ezc3d isEmpty(): negative residual word), else 0 with xyz = 0 like the node's targetPos_.zero_() (node.cpp:682-683)
extern "C" int smplpp_c3d_read(const smplpp_c3d_t * c, int64_t first, int64_t count, float * xyz, uint8_t * valid)
{
  if(!c || !xyz || !valid || first < 0 || count < 0 || first + count > c->frames)
    return fail(SMPLPP_ERR_INVALID, "C3D", "frame range outside the C3D file");
  const bool is_float = c->scale < 0.f;
  const size_t word = is_float ? 4 : 2;
  const size_t frame_bytes = (static_cast<size_t>(c->points) * 4 + c->analog_per_frame) * word;
  const float s = std::fabs(c->scale);
  for(int64_t f = 0; f < count; f++)
  {
    const size_t base = c->data_offset + static_cast<size_t>(first + f) * frame_bytes;
    for(int k = 0; k < c->points; k++)
    {
      float v[4];
      if(is_float)
        memcpy(v, c->bytes.data() + base + static_cast<size_t>(k) * 16, 16);
      else
      {
        int16_t w[4];
        memcpy(w, c->bytes.data() + base + static_cast<size_t>(k) * 8, 8);
        v[0] = w[0] * s, v[1] = w[1] * s, v[2] = w[2] * s, v[3] = w[3];
      }
      const bool ok = v[3] >= 0.f;
      const size_t o = (static_cast<size_t>(f) * c->points + k);
      valid[o] = ok ? 1 : 0;
      xyz[3 * o] = ok ? v[0] : 0.f, xyz[3 * o + 1] = ok ? v[1] : 0.f, xyz[3 * o + 2] = ok ? v[2] : 0.f;
    }
  }
  return SMPLPP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// result files of the mocap modes
// ------------------------------------------------------------------------------------------------------------
// MocapBody.yaml (node/node.cpp:1425-1441): the body stage's result, loaded back as ROS parameters for the motion
// stage (node/node.cpp:509-535).  Layout:
//   beta: [b0, ..., b9]
//   ikTaskList:
//     - name: <marker>
//       faceIdx: <int>
//       vertexWeights: [w0, w1, w2]
// The reference formats the numbers with Eigen::FullPrecision (Eigen is not under /root/reference, its digit count for
// float is version dependent); here every float is written with 9 significant digits, which round-trips exactly.
extern "C" int smplpp_write_mocap_body_yaml(const char * path, const float * beta10, int32_t n, const char * const * names,
                                            const int64_t * face_idx, const float * vertex_weights)
{
  if(!path || !beta10 || n < 0 || (n > 0 && (!names || !face_idx || !vertex_weights)))
    return fail(SMPLPP_ERR_INVALID, "node", "invalid mocap body description");
  FILE * f = fopen(path, "w");
  if(!f) return fail(SMPLPP_ERR_IO, "node", std::string("Cannot write ") + path);
  fprintf(f, "beta: [");
  for(int i = 0; i < kShapeDim; i++) fprintf(f, "%s%.9g", i ? ", " : "", static_cast<double>(beta10[i]));
  fprintf(f, "]\nikTaskList:\n");
  for(int i = 0; i < n; i++)
  {
    fprintf(f, "  - name: %s\n    faceIdx: %lld\n    vertexWeights: [%.9g, %.9g, %.9g]\n", names[i],
            static_cast<long long>(face_idx[i]), static_cast<double>(vertex_weights[3 * i]),
            static_cast<double>(vertex_weights[3 * i + 1]), static_cast<double>(vertex_weights[3 * i + 2]));
  }
  fclose(f);
  return SMPLPP_OK;
}

struct smplpp_mocap_body
{
  std::vector<float> beta;
  std::vector<std::string> names;
  std::vector<int64_t> face_idx;
  std::vector<float> weights;
};

namespace
{
bool parse_bracket_list(const std::string & line, size_t from, std::vector<float> & out)
{
  const size_t a = line.find('[', from), b = line.find(']', from);
  if(a == std::string::npos || b == std::string::npos || b < a) return false;
  const char * p = line.c_str() + a + 1;
  const char * end = line.c_str() + b;
  while(p < end)
  {
    char * stop = nullptr;
    const double v = strtod(p, &stop);
    if(stop == p) break;
    out.push_back(static_cast<float>(v));
    p = stop;
    while(p < end && (*p == ',' || *p == ' ')) p++;
  }
  return true;
}
} // namespace

extern "C" int smplpp_mocap_body_open(const char * path, smplpp_mocap_body_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "node", "null output");
  std::ifstream f(path ? path : "");
  if(!f) return fail(SMPLPP_ERR_IO, "node", std::string("Cannot open ") + (path ? path : ""));
  std::unique_ptr<smplpp_mocap_body> b(new smplpp_mocap_body());
  std::string line;
  while(std::getline(f, line))
  {
    size_t p;
    if(line.compare(0, 5, "beta:") == 0)
      parse_bracket_list(line, 5, b->beta);
    else if((p = line.find("- name:")) != std::string::npos)
    {
      std::string name = line.substr(p + 7);
      while(!name.empty() && name.front() == ' ') name.erase(name.begin());
      while(!name.empty() && (name.back() == ' ' || name.back() == '\r')) name.pop_back();
      b->names.push_back(name);
    }
    else if((p = line.find("faceIdx:")) != std::string::npos)
      b->face_idx.push_back(strtoll(line.c_str() + p + 8, nullptr, 10));
    else if((p = line.find("vertexWeights:")) != std::string::npos)
      parse_bracket_list(line, p, b->weights);
  }
  if(b->beta.size() != static_cast<size_t>(kShapeDim))
    return fail(SMPLPP_ERR_IO, "node", "Size of beta must be " + std::to_string(kShapeDim) + " but " + std::to_string(b->beta.size())); // node.cpp:511-515
  if(b->face_idx.size() != b->names.size() || b->weights.size() != 3 * b->names.size())
    return fail(SMPLPP_ERR_IO, "node", "malformed ikTaskList in the mocap body file");
  *out = b.release();
  return SMPLPP_OK;
}
extern "C" void smplpp_mocap_body_close(smplpp_mocap_body_t * b)
{
  delete b;
}
extern "C" int32_t smplpp_mocap_body_task_count(const smplpp_mocap_body_t * b)
{
  return b ? static_cast<int32_t>(b->names.size()) : 0;
}
extern "C" const char * smplpp_mocap_body_task_name(const smplpp_mocap_body_t * b, int32_t i)
{
  return (b && i >= 0 && i < static_cast<int32_t>(b->names.size())) ? b->names[static_cast<size_t>(i)].c_str() : "";
}
// beta10 (10), face_idx (n), vertex_weights (n, 3)
extern "C" int smplpp_mocap_body_get(const smplpp_mocap_body_t * b, float * beta10, int64_t * face_idx, float * vertex_weights)
{
  if(!b || !beta10 || !face_idx || !vertex_weights) return fail(SMPLPP_ERR_INVALID, "node", "null argument");
  std::copy(b->beta.begin(), b->beta.end(), beta10);
  std::copy(b->face_idx.begin(), b->face_idx.end(), face_idx);
  std::copy(b->weights.begin(), b->weights.end(), vertex_weights);
  return SMPLPP_OK;
}

// Motion as text: one line per frame, the 75 values of theta (25 x 3 row-major: translation, then 24 axis-angles)
// separated by blanks (scripts/convertRosbagToText.py:13-19 writes exactly this from the smplpp/motion message)
extern "C" int smplpp_write_motion_text(const char * path, int64_t frames, const float * theta75)
{
  if(!path || frames < 0 || (frames > 0 && !theta75)) return fail(SMPLPP_ERR_INVALID, "node", "invalid motion");
  FILE * f = fopen(path, "w");
  if(!f) return fail(SMPLPP_ERR_IO, "node", std::string("Cannot write ") + path);
  for(int64_t i = 0; i < frames; i++)
  {
    for(int k = 0; k < 75; k++) fprintf(f, "%s%.9g", k ? " " : "", static_cast<double>(theta75[i * 75 + k]));
    fputc('\n', f);
  }
  fclose(f);
  return SMPLPP_OK;
}

// frames_out = number of lines; theta75 may be null to query the count first; at most max_frames lines are stored
extern "C" int smplpp_read_motion_text(const char * path, int64_t max_frames, float * theta75, int64_t * frames_out)
{
  if(!frames_out) return fail(SMPLPP_ERR_INVALID, "node", "null argument");
  std::ifstream f(path ? path : "");
  if(!f) return fail(SMPLPP_ERR_IO, "node", std::string("Cannot open ") + (path ? path : ""));
  std::string line;
  int64_t n = 0;
  while(std::getline(f, line))
  {
    if(line.find_first_not_of(" \t\r") == std::string::npos) continue;
    const char * p = line.c_str();
    float v[75];
    int k = 0;
    for(; k < 75; k++)
    {
      char * stop = nullptr;
      v[k] = strtof(p, &stop);
      if(stop == p) break;
      p = stop;
    }
    if(k != 75) return fail(SMPLPP_ERR_IO, "node", "a motion line does not hold 75 values (line " + std::to_string(n + 1) + ")");
    if(theta75 && n < max_frames) std::copy(v, v + 75, theta75 + n * 75);
    n++;
  }
  *frames_out = n;
  return SMPLPP_OK;
}

// SMPL::out (src/SMPL.cpp:757-790): Wavefront OBJ of one mesh, "v x y z" per vertex in the default ostream float format
// (6 significant digits = printf %g) and "f a b c" per face with the stored 1-based indices
extern "C" int smplpp_write_obj(const char * path, int64_t n_vertices, const float * vertices, int64_t n_faces,
                                const int32_t * face_indices_1based)
{
  if(!path || n_vertices < 1 || !vertices || n_faces < 0 || (n_faces > 0 && !face_indices_1based))
    return fail(SMPLPP_ERR_INVALID, "SMPL", "Cannot export the deformed mesh!"); // SMPL.cpp:785
  FILE * f = fopen(path, "w");
  if(!f) return fail(SMPLPP_ERR_IO, "SMPL", "Cannot export the deformed mesh!");
  for(int64_t i = 0; i < n_vertices; i++)
    fprintf(f, "v %g %g %g\n", static_cast<double>(vertices[3 * i]), static_cast<double>(vertices[3 * i + 1]),
            static_cast<double>(vertices[3 * i + 2]));
  for(int64_t i = 0; i < n_faces; i++)
    fprintf(f, "f %d %d %d\n", face_indices_1based[3 * i], face_indices_1based[3 * i + 1], face_indices_1based[3 * i + 2]);
  fclose(f);
  return SMPLPP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// .npz twin of the model file (scripts/preprocess.py:98-117 writes it with np.savez next to the JSON): a ZIP archive of
// STORED .npy members.  Read through the central directory (numpy writes zip64 extra fields), C-ordered little-endian
// f4 / f8 / i4 / i8 / u4 / u8 arrays.  Values are widened to double like the JSON reader's, so both loaders share the
// model assembly below.
// ------------------------------------------------------------------------------------------------------------
namespace
{
bool npz_read(const std::string & b, smplpp_json & out, std::string & err)
{
  // end of central directory: scan back for its signature
  if(b.size() < 22) return err = "not a zip archive", false;
  size_t eocd = std::string::npos;
  for(size_t i = b.size() - 22 + 1; i-- > 0;)
  {
    if(rd<uint32_t>(b, i) == 0x06054b50u)
    {
      eocd = i;
      break;
    }
    if(b.size() - i > 22 + 65535) break;
  }
  if(eocd == std::string::npos) return err = "not a zip archive (no end-of-central-directory record)", false;
  uint64_t n_entries = rd<uint16_t>(b, eocd + 10), cd_off = rd<uint32_t>(b, eocd + 16);
  if(cd_off == 0xFFFFFFFFu || n_entries == 0xFFFFu)
  {
    // zip64: locator 20 bytes before the EOCD points at the zip64 EOCD record
    if(eocd < 20 || rd<uint32_t>(b, eocd - 20) != 0x07064b50u) return err = "zip64 locator missing", false;
    const uint64_t z = rd<uint64_t>(b, eocd - 20 + 8);
    if(z + 56 > b.size() || rd<uint32_t>(b, z) != 0x06064b50u) return err = "zip64 end-of-central-directory record missing", false;
    n_entries = rd<uint64_t>(b, z + 32);
    cd_off = rd<uint64_t>(b, z + 48);
  }
  size_t pos = cd_off;
  for(uint64_t e = 0; e < n_entries; e++)
  {
    if(pos + 46 > b.size() || rd<uint32_t>(b, pos) != 0x02014b50u) return err = "corrupt central directory", false;
    const int method = rd<uint16_t>(b, pos + 10);
    uint64_t csize = rd<uint32_t>(b, pos + 20), usize = rd<uint32_t>(b, pos + 24), lho = rd<uint32_t>(b, pos + 42);
    const int nlen = rd<uint16_t>(b, pos + 28), xlen = rd<uint16_t>(b, pos + 30), clen = rd<uint16_t>(b, pos + 32);
    // every offset below comes from the archive itself: nothing is read before it is checked against the file size
    if(pos + 46 + static_cast<size_t>(nlen) + xlen + clen > b.size()) return err = "corrupt central directory entry", false;
    std::string name(b.data() + pos + 46, nlen);
    // zip64 extended information (header id 1): the fields that are 0xFFFFFFFF above, in order
    size_t x = pos + 46 + nlen;
    const size_t xend = x + xlen;
    while(x + 4 <= xend)
    {
      const int id = rd<uint16_t>(b, x), sz = rd<uint16_t>(b, x + 2);
      if(x + 4 + static_cast<size_t>(sz) > xend) return err = "corrupt extra field of " + name, false;
      if(id == 1)
      {
        size_t q = x + 4;
        const size_t qend = x + 4 + sz;
        auto take64 = [&](uint64_t & dst) -> bool {
          if(q + 8 > qend) return false;
          dst = rd<uint64_t>(b, q), q += 8;
          return true;
        };
        if(usize == 0xFFFFFFFFu && !take64(usize)) return err = "truncated zip64 field of " + name, false;
        if(csize == 0xFFFFFFFFu && !take64(csize)) return err = "truncated zip64 field of " + name, false;
        if(lho == 0xFFFFFFFFu && !take64(lho)) return err = "truncated zip64 field of " + name, false;
      }
      x += 4 + sz;
    }
    pos = xend + clen;
    if(name.size() < 4 || name.compare(name.size() - 4, 4, ".npy") != 0) continue;
    if(method != 0) return err = "compressed .npz members are not supported (np.savez writes them stored): " + name, false;
    if(lho > b.size() || b.size() - lho < 30 || rd<uint32_t>(b, lho) != 0x04034b50u)
      return err = "corrupt local header of " + name, false;
    const size_t data = lho + 30 + rd<uint16_t>(b, lho + 26) + rd<uint16_t>(b, lho + 28);
    if(data > b.size() || usize > b.size() - data || usize < 12) return err = "truncated member " + name, false;
    // .npy: magic, version, header length, python dict literal
    if(memcmp(b.data() + data, "\x93NUMPY", 6) != 0) return err = "not a .npy member: " + name, false;
    const int major = static_cast<uint8_t>(b[data + 6]);
    const size_t hlen = major >= 2 ? rd<uint32_t>(b, data + 8) : rd<uint16_t>(b, data + 8);
    const size_t hoff = data + (major >= 2 ? 12 : 10);
    if(hlen > usize || hoff + hlen > data + usize) return err = "truncated .npy header of " + name, false;
    const std::string hdr(b.data() + hoff, hlen);
    constexpr size_t npos = std::string::npos;
    auto after = [&](const char * key) -> size_t {
      const size_t k = hdr.find(key);
      if(k == npos) return npos;
      const size_t c = hdr.find(':', k);
      return c == npos ? npos : c + 1;
    };
    size_t p = after("'descr'");
    if(p == npos) return err = "no descr in " + name, false;
    const size_t q1 = hdr.find('\'', p);
    const size_t q2 = q1 == npos ? npos : hdr.find('\'', q1 + 1);
    if(q2 == npos) return err = "malformed descr in " + name, false;
    const std::string descr = hdr.substr(q1 + 1, q2 - q1 - 1);
    if(descr.size() < 3) return err = "unsupported dtype " + descr + " of " + name, false;
    p = after("'fortran_order'");
    if(p != npos)
    {
      const size_t v = hdr.find_first_not_of(' ', p);
      if(v != npos && hdr.compare(v, 4, "True") == 0) return err = "Fortran-ordered array " + name + " is not supported", false;
    }
    p = after("'shape'");
    const size_t s1 = p == npos ? npos : hdr.find('(', p);
    const size_t s2 = s1 == npos ? npos : hdr.find(')', s1);
    if(s2 == npos) return err = "no shape in " + name, false;
    smplpp_json::Array a;
    {
      const char * c = hdr.c_str() + s1 + 1;
      const char * cend = hdr.c_str() + s2;
      while(c < cend)
      {
        char * stop = nullptr;
        const long long v = strtoll(c, &stop, 10);
        if(stop == c) break;
        if(v < 0) return err = "negative dimension in " + name, false;
        a.shape.push_back(v);
        c = stop;
        while(c < cend && (*c == ',' || *c == ' ')) c++;
      }
    }
    const size_t payload = hoff + hlen;
    const size_t esz = static_cast<size_t>(descr[2] - '0');
    if((descr[0] != '<' && descr[0] != '|') || (esz != 4 && esz != 8) || descr.size() != 3)
      return err = "unsupported dtype " + descr + " of " + name, false;
    const size_t avail = (data + usize - payload) / esz; // elements the member can hold: the product may not exceed it
    size_t count = 1;
    for(int64_t d : a.shape)
    {
      if(d != 0 && count > avail / static_cast<size_t>(d)) return err = "shape of " + name + " exceeds the member size", false;
      count *= static_cast<size_t>(d);
    }
    if(count > avail) return err = "shape of " + name + " exceeds the member size", false;
    a.data.resize(count);
    for(size_t i = 0; i < count; i++)
    {
      const size_t o = payload + i * esz;
      double v = 0.0;
      if(descr[1] == 'f') v = esz == 4 ? static_cast<double>(rd<float>(b, o)) : rd<double>(b, o);
      else if(descr[1] == 'i') v = esz == 4 ? static_cast<double>(rd<int32_t>(b, o)) : static_cast<double>(rd<int64_t>(b, o));
      else if(descr[1] == 'u') v = esz == 4 ? static_cast<double>(rd<uint32_t>(b, o)) : static_cast<double>(rd<uint64_t>(b, o));
      else return err = "unsupported dtype " + descr + " of " + name, false;
      a.data[i] = v;
    }
    out.arrays[name.substr(0, name.size() - 4)] = std::move(a);
  }
  return true;
}
} // namespace

// the same handle type and accessor (smplpp_json_array) as the JSON reader
extern "C" int smplpp_npz_open(const char * path, smplpp_json_t ** out)
{
  if(!out) return fail(SMPLPP_ERR_INVALID, "NPZ", "null output");
  std::string bytes, err;
  if(!path || !read_file(path, bytes)) return fail(SMPLPP_ERR_IO, "NPZ", "Cannot find the .npz file!");
  std::unique_ptr<smplpp_json> j(new smplpp_json());
  try
  {
    if(!npz_read(bytes, *j, err)) return fail(SMPLPP_ERR_IO, "NPZ", "Cannot read the .npz file: " + err);
  }
  catch(const std::exception & ex) // no C++ exception may cross the C boundary
  {
    return fail(SMPLPP_ERR_IO, "NPZ", std::string("Cannot read the .npz file: ") + ex.what());
  }
  *out = j.release();
  return SMPLPP_OK;
}
