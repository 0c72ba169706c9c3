// psdr-b200 host layer: the pybind11 module behind `import psdr_cuda` — C++ host code over the C ABI of
// include/psdr_b200.h, mirroring the Python-visible surface of the reference's module (src/psdr.cpp:41-295):
// RenderOption, Scene (load_file / load_string / configure / opts / num_sensors / num_meshes / param_map), Mesh,
// DiffuseBSDF, RoughConductorBSDF, Bitmap1fD / Bitmap3fD, PerspectiveCamera, AreaLight, EnvironmentMap and the
// integrators (DirectIntegrator, FieldExtractionIntegrator, + PathIntegrator) with renderC / renderD.
//
// Differences that are deliberate: arrays cross the boundary as numpy arrays / raw device pointers instead of Enoki
// arrays (Enoki is not a dependency; psdr_cuda/__init__.py wraps device results as torch tensors and wires renderD into
// torch.autograd); scene ingest (XML subset + OBJ) is implemented here without pugixml / tinyobj.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <unistd.h>

#include <array>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "../../include/psdr_b200.h"
#include "pb_xml.h"

namespace py = pybind11;
using namespace pbhost;
using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;
using iarray = py::array_t<int, py::array::c_style | py::array::forcecast>;

namespace {

struct Mat4 {
    float m[16];
    Mat4() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
};
Mat4 operator*(const Mat4 &a, const Mat4 &b) {
    Mat4 c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s = s + a.m[4 * i + k] * b.m[4 * k + j];
            c.m[4 * i + j] = s;
        }
    return c;
}
farray mat_to_numpy(const Mat4 &m) {
    farray a({4, 4});
    std::memcpy(a.mutable_data(), m.m, sizeof(m.m));
    return a;
}
Mat4 mat_from_numpy(const farray &a) {
    if (a.ndim() != 2 || a.shape(0) != 4 || a.shape(1) != 4) throw std::runtime_error("expected a 4x4 matrix");
    Mat4 m;
    std::memcpy(m.m, a.data(), sizeof(m.m));
    return m;
}

// ---- transform.h:14-79 ---------------------------------------------------------------------------------------------
Mat4 m_translate(float x, float y, float z) { Mat4 m; m.m[3] = x; m.m[7] = y; m.m[11] = z; return m; }
Mat4 m_scale(float x, float y, float z) { Mat4 m; m.m[0] = x; m.m[5] = y; m.m[10] = z; return m; }
Mat4 m_rotate(float x, float y, float z, float angle_deg) {
    const float ang = angle_deg * (3.14159265358979323846f / 180.f);   // transform.h:27: deg_to_rad(a) = a * (Pi / 180)
    const float s = (float)std::sin((double)ang), c = (float)std::cos((double)ang), k = 1.f - c;
    Mat4 m;
    m.m[0] = x * x * k + c;     m.m[1] = x * y * k - z * s; m.m[2] = x * z * k + y * s;
    m.m[4] = y * x * k + z * s; m.m[5] = y * y * k + c;     m.m[6] = y * z * k - x * s;
    m.m[8] = z * x * k - y * s; m.m[9] = z * y * k + x * s; m.m[10] = z * z * k + c;
    return m;
}
void norm3(float *v) { const float n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int k = 0; k < 3; ++k) v[k] /= n; }
void cross3(const float *a, const float *b, float *o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
Mat4 m_look_at(const float *origin, const float *target, const float *up) {
    float dir[3] = {target[0] - origin[0], target[1] - origin[1], target[2] - origin[2]}, left[3], new_up[3];
    norm3(dir);
    cross3(up, dir, left); norm3(left);
    cross3(dir, left, new_up);
    Mat4 m;
    for (int i = 0; i < 3; ++i) { m.m[4 * i] = left[i]; m.m[4 * i + 1] = new_up[i]; m.m[4 * i + 2] = dir[i]; m.m[4 * i + 3] = origin[i]; }
    return m;
}

std::vector<float> parse_vector(const std::string &str, size_t n, bool allow_short = false) {   // scene_loader.cpp:23-52
    std::vector<float> v;
    std::string tok;
    for (size_t i = 0; i <= str.size(); ++i) {
        const char ch = i < str.size() ? str[i] : ',';
        if (ch == ',' || ch == ' ' || ch == '\t' || ch == '\n') { if (!tok.empty()) { v.push_back((float)std::atof(tok.c_str())); tok.clear(); } }
        else tok.push_back(ch);
    }
    if (v.size() > n) throw std::runtime_error("Vector too long: [" + str + "]");
    if (v.size() < n) {
        if (!allow_short) throw std::runtime_error("Vector too short: [" + str + "]");
        const float fill = v.empty() ? 0.f : v.back();
        v.resize(n, fill);
    }
    return v;
}

Mat4 load_transform(const XmlNode *node) {   // scene_loader.cpp:80-127: left-multiplied in document order
    Mat4 result;
    if (!node) return result;
    const std::string nm = node->attr("name");
    if (nm != "to_world" && nm != "toWorld") throw std::runtime_error("Invalid transformation name: " + nm);
    for (auto &op : node->children) {
        Mat4 t;
        if (op->name == "translate") t = m_translate(op->attr_float("x", 0.f), op->attr_float("y", 0.f), op->attr_float("z", 0.f));
        else if (op->name == "rotate") t = m_rotate(op->attr_float("x", 0.f), op->attr_float("y", 0.f), op->attr_float("z", 0.f), op->attr_float("angle", 0.f));
        else if (op->name == "scale") t = m_scale(op->attr_float("x", 1.f), op->attr_float("y", 1.f), op->attr_float("z", 1.f));
        else if (op->name == "look_at" || op->name == "lookAt" || op->name == "lookat") {
            auto o = parse_vector(op->attr("origin"), 3), tg = parse_vector(op->attr("target"), 3), up = parse_vector(op->attr("up"), 3);
            t = m_look_at(o.data(), tg.data(), up.data());
        } else if (op->name == "matrix") {
            auto v = parse_vector(op->attr("value"), 16);
            std::memcpy(t.m, v.data(), sizeof(t.m));
        } else throw std::runtime_error("Unsupported transformation: " + op->name);
        result = t * result;
    }
    return result;
}

// ---- Python-visible objects ------------------------------------------------------------------------------------------
struct Object {
    std::string id;
    virtual ~Object() = default;
    virtual std::string type_name() const = 0;
    virtual std::string to_string() const { return type_name() + (id.empty() ? "" : "[id=" + id + "]"); }
};

struct Bitmap {   // Bitmap1fD / Bitmap3fD (src/core/bitmap.cpp, src/psdr.cpp:102-120)
    int channels = 3, width = 1, height = 1;
    std::vector<float> data;   // interleaved
    bool dirty = true, requires_grad = false;
    explicit Bitmap(int c = 3, float v = 0.f) : channels(c), data(c, v) {}
    void fill(const std::vector<float> &v) { width = height = 1; data = v; data.resize(channels, v.empty() ? 0.f : v.back()); dirty = true; }
    farray get_data() const {
        farray a(std::vector<py::ssize_t>{(py::ssize_t)width * height, (py::ssize_t)channels});
        std::memcpy(a.mutable_data(), data.data(), data.size() * sizeof(float));
        return a;
    }
    void set_data(const farray &a) {
        if ((size_t)a.size() != (size_t)width * height * channels) {
            if (a.size() == channels) { width = height = 1; }
            else throw std::runtime_error("Bitmap.data: size does not match resolution (set resolution first)");
        }
        data.assign(a.data(), a.data() + a.size());
        dirty = true;
    }
    // `bsdf.reflectance = Bitmap3fD(...)`: the reference's members are read-write (src/psdr.cpp:208-215,236); the value is copied into the
    // scene's own bitmap, which keeps its place in the gradient vector
    void assign(const Bitmap &o) {
        if (o.channels != channels) throw std::runtime_error("Bitmap: expected " + std::to_string(channels) + " channel(s), got " + std::to_string(o.channels));
        width = o.width; height = o.height; data = o.data; dirty = true;
    }
};

struct BSDF : Object { int index = -1; };
struct Diffuse : BSDF {
    Bitmap reflectance{3, .5f};
    std::string type_name() const override { return "Diffuse"; }                // the C++ class name, as PSDR_CLASS_DECL_END yields (diffuse.h:41)
};
struct RoughConductor : BSDF {
    Bitmap alpha_u{1, .1f}, alpha_v{1, .1f}, eta{3, 0.f}, k{3, 1.f}, specular_reflectance{3, 1.f};
    std::string type_name() const override { return "RoughConductor"; }         // roughconductor.h:54
};

struct Mesh : Object {
    int index = -1, bsdf = -1, emitter = -1;
    std::vector<float> verts, uvs;
    std::vector<int> faces, uv_faces;
    bool use_face_normals = false, enable_edges = true, requires_grad = false, uv_requires_grad = false;
    bool verts_dirty = false, transform_dirty = false, uvs_dirty = false, uploaded = false;
    Mat4 to_world_raw, to_world_left, to_world_right;
    int nv() const { return (int)verts.size() / 3; }
    int nf() const { return (int)faces.size() / 3; }
    std::string type_name() const override { return "Mesh"; }

    void load(const std::string &path) {   // Mesh::load (mesh.cpp:62-141): positions, texcoords, fan-triangulated faces
        std::ifstream in(path);
        if (!in) throw std::runtime_error("Failed to load OBJ from: " + path);
        verts.clear(); uvs.clear(); faces.clear(); uv_faces.clear();
        std::string line;
        std::vector<int> cv, ct;
        while (std::getline(in, line)) {
            if (line.size() < 2) continue;
            if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
                const char *p = line.c_str() + 1; char *e;
                for (int k = 0; k < 3; ++k) { verts.push_back(std::strtof(p, &e)); p = e; }
            } else if (line[0] == 'v' && line[1] == 't') {
                const char *p = line.c_str() + 2; char *e;
                for (int k = 0; k < 2; ++k) { uvs.push_back(std::strtof(p, &e)); p = e; }
            } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
                cv.clear(); ct.clear();
                std::istringstream ss(line.substr(1));
                std::string tok;
                while (ss >> tok) {
                    int vi = 0, ti = 0;
                    const size_t s1 = tok.find('/');
                    vi = std::atoi(tok.substr(0, s1).c_str());
                    if (s1 != std::string::npos) {
                        const size_t s2 = tok.find('/', s1 + 1);
                        const std::string t = tok.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                        if (!t.empty()) ti = std::atoi(t.c_str());
                    }
                    cv.push_back(vi > 0 ? vi - 1 : nv() + vi);
                    ct.push_back(ti > 0 ? ti - 1 : (ti < 0 ? (int)uvs.size() / 2 + ti : -1));
                }
                for (size_t k = 2; k < cv.size(); ++k) {
                    faces.insert(faces.end(), {cv[0], cv[k - 1], cv[k]});
                    uv_faces.insert(uv_faces.end(), {ct[0], ct[k - 1], ct[k]});
                }
            }
        }
        if (verts.empty()) throw std::runtime_error("Failed to load OBJ from: " + path);
        if (uvs.empty()) uv_faces.clear();
    }
    farray get_vertices() const {
        farray a(std::vector<py::ssize_t>{(py::ssize_t)nv(), 3});
        std::memcpy(a.mutable_data(), verts.data(), verts.size() * sizeof(float));
        return a;
    }
    void set_vertices(const farray &a) {
        if ((size_t)a.size() != verts.size()) throw std::runtime_error("vertex_positions: wrong size");
        verts.assign(a.data(), a.data() + a.size());
        verts_dirty = true;
    }
    farray get_uvs() const {
        farray a(std::vector<py::ssize_t>{(py::ssize_t)(uvs.size() / 2), 2});
        std::memcpy(a.mutable_data(), uvs.data(), uvs.size() * sizeof(float));
        return a;
    }
    void set_uvs(const farray &a) {
        if ((size_t)a.size() != uvs.size()) throw std::runtime_error("vertex_uv: wrong size");
        uvs.assign(a.data(), a.data() + a.size());
        uvs_dirty = true;
    }
    iarray get_faces() const {
        iarray a(std::vector<py::ssize_t>{(py::ssize_t)nf(), 3});
        std::memcpy(a.mutable_data(), faces.data(), faces.size() * sizeof(int));
        return a;
    }
    // face_indices / face_uv_indices are read-write in the reference (src/psdr.cpp:255-256). The topology goes to the device with the first
    // Scene.configure (BVH, edge lists), so it can be edited until then.
    void set_index_array(std::vector<int> &dst, const iarray &a, int limit, const char *what) {
        if (uploaded) throw std::runtime_error(std::string(what) + ": the topology of a mesh that is already on the device cannot be edited");
        if (a.ndim() != 2 || a.shape(1) != 3) throw std::runtime_error(std::string(what) + ": expected an (n, 3) integer array");
        for (py::ssize_t i = 0; i < a.size(); ++i) if (a.data()[i] < 0 || a.data()[i] >= limit) throw std::runtime_error(std::string(what) + ": index out of range");
        dst.assign(a.data(), a.data() + a.size());
    }
    void set_faces(const iarray &a) { set_index_array(faces, a, nv(), "face_indices"); }
    void set_uv_faces(const iarray &a) {
        if ((size_t)a.size() != faces.size()) throw std::runtime_error("face_uv_indices: one uv triple per face");
        set_index_array(uv_faces, a, (int)uvs.size() / 2, "face_uv_indices");
    }
    // object-space, area-weighted vertex normals (mesh.cpp:19-51 on the raw positions, mesh.cpp:219)
    std::vector<float> vertex_normals() const {
        std::vector<float> n(verts.size(), 0.f), w(nv(), 0.f);
        for (int f = 0; f < nf(); ++f) {
            const float *a = &verts[3 * faces[3 * f]], *b = &verts[3 * faces[3 * f + 1]], *c = &verts[3 * faces[3 * f + 2]];
            const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            const float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            const float area = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
            for (int k = 0; k < 3; ++k) {
                const int v = faces[3 * f + k];
                for (int d = 0; d < 3; ++d) n[3 * v + d] += fn[d];
                w[v] += area;
            }
        }
        for (int v = 0; v < nv(); ++v) {
            float x[3] = {n[3 * v] / w[v], n[3 * v + 1] / w[v], n[3 * v + 2] / w[v]};
            const float inv = 1.f / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            for (int d = 0; d < 3; ++d) n[3 * v + d] = x[d] * inv;
        }
        return n;
    }
    // Mesh::dump (mesh.cpp:318-392): object-space OBJ in the reference's layout — "v" (+ "vn" per vertex unless face normals are used)
    // in %.6e, then "vt", then faces as v, v/vt, v//vn or v/vt/vn
    void dump(const std::string &fname) const {
        FILE *out = std::fopen(fname.c_str(), "wt");
        if (!out) throw std::runtime_error("Failed to open: " + fname);
        std::vector<float> vn;
        if (!use_face_normals) vn = vertex_normals();
        for (int v = 0; v < nv(); ++v) {
            std::fprintf(out, "v %.6e %.6e %.6e\n", verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]);
            if (!use_face_normals) std::fprintf(out, "vn %.6e %.6e %.6e\n", vn[3 * v], vn[3 * v + 1], vn[3 * v + 2]);
        }
        const bool has_uv = !uvs.empty();
        for (size_t t = 0; t + 1 < uvs.size(); t += 2) std::fprintf(out, "vt %.6e %.6e\n", (double)uvs[t], (double)uvs[t + 1]);
        for (int f = 0; f < nf(); ++f) {
            const int v0 = faces[3 * f] + 1, v1 = faces[3 * f + 1] + 1, v2 = faces[3 * f + 2] + 1;
            if (has_uv) {
                const int t0 = uv_faces[3 * f] + 1, t1 = uv_faces[3 * f + 1] + 1, t2 = uv_faces[3 * f + 2] + 1;
                if (use_face_normals) std::fprintf(out, "f %d/%d %d/%d %d/%d\n", v0, t0, v1, t1, v2, t2);
                else std::fprintf(out, "f %d/%d/%d %d/%d/%d %d/%d/%d\n", v0, t0, v0, v1, t1, v1, v2, t2, v2);
            } else if (use_face_normals) std::fprintf(out, "f %d %d %d\n", v0, v1, v2);
            else std::fprintf(out, "f %d//%d %d//%d %d//%d\n", v0, v0, v1, v1, v2, v2);
        }
        std::fclose(out);
    }
};

struct Sensor : Object {
    float fov_x = 45.f, near_clip = 0.1f, far_clip = 1e4f;
    Mat4 to_world;
    bool dirty = false, requires_grad = false;
    std::string type_name() const override { return "PerspectiveCamera"; }
};
struct Emitter : Object {};
struct AreaLight : Emitter {
    float radiance[3] = {0, 0, 0};
    int mesh = -1;
    std::string type_name() const override { return "AreaLight"; }
};
struct EnvironmentMap : Emitter {
    Bitmap radiance{3, 0.f};
    float scale = 1.f;
    bool scale_dirty = false, scale_requires_grad = false, transform_requires_grad = false;
    Mat4 to_world_raw, to_world_left;
    bool transform_dirty = false;
    std::string type_name() const override { return "AreaLight"; }   // sic: envmap.h:58
    std::string to_string() const override { return "EnvironmentMap[sampling_weight = 1]"; }   // envmap.cpp:146-150 (weight of a lone emitter)
};

struct RenderOption {   // types.h:171-182, psdr.cpp:53-72
    int width = 0, height = 0, spp = 0, sppe = 0, sppse = 0, log_level = 1;
    RenderOption() = default;
    RenderOption(int w, int h, int s) : width(w), height(h), spp(s), sppe(s), sppse(s) {}
    RenderOption(int w, int h, int s, int se) : width(w), height(h), spp(s), sppe(se), sppse(se) {}
    RenderOption(int w, int h, int s, int se, int sse) : width(w), height(h), spp(s), sppe(se), sppse(sse) {}
};

std::string find_file(const std::string &name, const std::string &xml_dir) {
    std::vector<std::string> trial = {name};
    std::string d = xml_dir;
    for (int up = 0; up < 5 && !d.empty(); ++up) {
        trial.push_back(d + "/" + name);
        const size_t s = d.find_last_of('/');
        if (s == std::string::npos || s == 0) break;
        d = d.substr(0, s);
    }
    for (auto &t : trial) { std::ifstream f(t); if (f.good()) return t; }
    throw std::runtime_error("Failed to load file: " + name);
}

class Scene {
public:
    pb_ctx *ctx = nullptr;
    RenderOption opts;
    std::vector<std::shared_ptr<Sensor>> sensors;
    std::vector<std::shared_ptr<BSDF>> bsdfs;
    std::vector<std::shared_ptr<Mesh>> meshes;
    std::vector<std::shared_ptr<Emitter>> emitters;
    std::shared_ptr<EnvironmentMap> envmap;
    std::map<std::string, std::shared_ptr<Object>> param_map;
    bool loaded = false, uploaded = false, configured = false;
    int shard_world = 1;   // > 1: the scene renders one shard of a multi-GPU job (set_shard / dist_init)

    // device < 0: description only (ingest and inspection work, configure / render need a context)
    explicit Scene(int device = 0) { if (device >= 0 && pb_ctx_create(device, &ctx)) throw std::runtime_error(pb_last_error(nullptr)); }
    ~Scene() { if (ctx) pb_ctx_destroy(ctx); }
    Scene(const Scene &) = delete;
    void check(int rc) const { if (rc) throw std::runtime_error(pb_last_error(ctx)); }
    int check_id(int rc) const { if (rc < 0) throw std::runtime_error(pb_last_error(ctx)); return rc; }

    void load_file(const std::string &fname, bool auto_configure) {
        std::ifstream in(fname);
        if (!in) throw std::runtime_error("XML parsing failed");
        std::stringstream ss;
        ss << in.rdbuf();
        std::string dir = ".";
        const size_t s = fname.find_last_of('/');
        if (s != std::string::npos) dir = fname.substr(0, s);
        if (!dir.empty() && dir[0] != '/') {
            char buf[4096];
            if (getcwd(buf, sizeof(buf))) dir = std::string(buf) + "/" + dir;
        }
        load_xml(ss.str(), dir, auto_configure);
    }
    void load_string(const std::string &xml, bool auto_configure) { load_xml(xml, "", auto_configure); }

    void load_texture(const XmlNode *node, Bitmap &bm, const std::string &xml_dir) {   // scene_loader.cpp:158-170
        if (node->name == "texture") {
            if (node->attr("type") != "bitmap") throw std::runtime_error("Unsupported texture type: " + node->attr("type"));
            const XmlNode *fn = node->child("string");
            if (!fn || fn->attr("name") != "filename") throw std::runtime_error("Failed to retrieve bitmap filename");
            load_exr(find_file(fn->attr("value"), xml_dir), bm);
        } else if (bm.channels == 1) {
            bm.fill({std::stof(node->attr("value"))});
        } else if (node->name == "float") {
            const float v = std::stof(node->attr("value"));
            bm.fill({v, v, v});
        } else if (node->name == "rgb") {
            bm.fill(parse_vector(node->attr("value"), 3, true));
        } else throw std::runtime_error("Unsupported RGB type: " + node->name);
    }
    static void load_exr(const std::string &path, Bitmap &bm) {   // bitmap_loader.cpp:13-53; decoding is delegated to Python (cv2)
        py::object img = py::module_::import("psdr_cuda").attr("_read_exr")(path, bm.channels);
        farray a = img.cast<farray>();
        bm.height = (int)a.shape(0); bm.width = (int)a.shape(1);
        bm.data.assign(a.data(), a.data() + a.size());
        bm.dirty = true;
    }

    void load_xml(const std::string &text, const std::string &xml_dir, bool auto_configure) {   // scene_loader.cpp:208-242
        if (loaded) throw std::runtime_error("Scene already loaded!");
        XmlParser parser(text);
        std::unique_ptr<XmlNode> doc = parser.parse();
        const XmlNode *root = doc->child("scene");
        if (!root) throw std::runtime_error("XML parsing failed");
        for (const XmlNode *node : root->all("sensor")) {
            const XmlNode *film = node->child("film"), *sampler = node->child("sampler");
            if (sensors.empty()) {
                if (!film) throw std::runtime_error("Missing film node");
                if (!sampler) throw std::runtime_error("Missing sampler node");
                const XmlNode *w = film->named({"width"}), *h = film->named({"height"});
                if (!w) throw std::runtime_error("Missing child node: width");
                if (!h) throw std::runtime_error("Missing child node: height");
                opts.width = std::stoi(w->attr("value")); opts.height = std::stoi(h->attr("value"));
                const XmlNode *cnt = sampler->child("integer");
                opts.spp = opts.sppe = opts.sppse = cnt ? std::stoi(cnt->attr("value")) : 0;
            } else {
                if (film) throw std::runtime_error("Duplicate film node");
                if (sampler) throw std::runtime_error("Duplicate sampler node");
            }
            if (node->attr("type") != "perspective") throw std::runtime_error("Unsupported sensor: " + node->attr("type"));
            auto s = std::make_shared<Sensor>();
            s->to_world = load_transform(node->child("transform"));
            const XmlNode *fov = node->named({"fov"});
            if (!fov) throw std::runtime_error("Missing child node: fov");
            s->fov_x = std::stof(fov->attr("value"));
            if (const XmlNode *ax = node->named({"fov_axis", "fovAxis"})) if (ax->attr("value") != "x") throw std::runtime_error("Unsupported fov-axis: " + ax->attr("value"));
            if (const XmlNode *n = node->named({"near_clip", "nearClip"})) s->near_clip = n->attr_float("value", 0.1f);
            if (const XmlNode *n = node->named({"far_clip", "farClip"})) s->far_clip = n->attr_float("value", 1e4f);
            sensors.push_back(s);
        }
        for (const XmlNode *node : root->all("bsdf")) {
            const std::string id = node->attr("id"), type = node->attr("type");
            if (id.empty()) throw std::runtime_error("BSDF must have an id");
            std::shared_ptr<BSDF> b;
            if (type == "diffuse") {
                auto d = std::make_shared<Diffuse>();
                const XmlNode *r = node->named({"reflectance"});
                if (!r) throw std::runtime_error("Missing child node: reflectance");
                load_texture(r, d->reflectance, xml_dir);
                b = d;
            } else if (type == "roughconductor") {
                auto r = std::make_shared<RoughConductor>();
                const XmlNode *alpha = node->named({"alpha"}), *eta = node->named({"eta"}), *k = node->named({"k"});
                if (!alpha) throw std::runtime_error("Missing child node: alpha");
                if (!eta) throw std::runtime_error("Missing child node: eta");
                if (!k) throw std::runtime_error("Missing child node: k");
                load_texture(alpha, r->alpha_u, xml_dir); load_texture(alpha, r->alpha_v, xml_dir);
                load_texture(eta, r->eta, xml_dir); load_texture(k, r->k, xml_dir);
                b = r;
            } else throw std::runtime_error("Unsupported BSDF: " + type);
            b->id = id; b->index = (int)bsdfs.size();
            if (param_map.count("BSDF[id=" + id + "]")) throw std::runtime_error("Duplicate BSDF id: " + id);
            param_map["BSDF[" + std::to_string(bsdfs.size()) + "]"] = b;
            param_map["BSDF[id=" + id + "]"] = b;
            bsdfs.push_back(b);
        }
        for (const XmlNode *node : root->all("emitter")) {
            if (node->attr("type") != "envmap") throw std::runtime_error("Unsupported emitter: " + node->attr("type"));
            if (envmap) throw std::runtime_error("A scene is only allowed to have one envmap!");
            const XmlNode *fn = node->child("string");
            if (!fn || fn->attr("name") != "filename") throw std::runtime_error("Failed to retrieve bitmap filename");
            auto e = std::make_shared<EnvironmentMap>();
            load_exr(find_file(fn->attr("value"), xml_dir), e->radiance);
            if (const XmlNode *sc = node->named({"scale"})) e->scale = sc->attr_float("value", 1.f);
            e->to_world_raw = load_transform(node->child("transform"));
            envmap = e;
            emitters.push_back(e);
        }
        for (const XmlNode *node : root->all("shape")) {
            if (node->attr("type") != "obj") throw std::runtime_error("Unsupported shape: " + node->attr("type"));
            const XmlNode *fn = node->child("string");
            if (!fn || fn->attr("name") != "filename") throw std::runtime_error("Missing mesh filename");
            auto m = std::make_shared<Mesh>();
            m->load(find_file(fn->attr("value"), xml_dir));
            const XmlNode *ref = node->child("ref");
            if (!ref) throw std::runtime_error("Missing BSDF reference");
            if (node->child("bsdf")) throw std::runtime_error("BSDFs declared under shapes are not supported.");
            auto it = param_map.find("BSDF[id=" + ref->attr("id") + "]");
            if (it == param_map.end()) throw std::runtime_error("Unknown BSDF id: " + ref->attr("id"));
            m->bsdf = std::static_pointer_cast<BSDF>(it->second)->index;
            if (const XmlNode *f = node->named({"face_normals", "faceNormals"})) m->use_face_normals = (f->attr("value") == "true");
            m->id = node->attr("id");
            m->to_world_raw = load_transform(node->child("transform"));
            m->index = (int)meshes.size();
            if (const XmlNode *em = node->child("emitter")) {
                if (em->attr("type") != "area") throw std::runtime_error("Unsupported emitter: " + em->attr("type"));
                const XmlNode *rad = em->named({"radiance"});
                if (!rad) throw std::runtime_error("Missing child node: radiance");
                auto a = std::make_shared<AreaLight>();
                std::vector<float> v = rad->name == "float" ? std::vector<float>(3, std::stof(rad->attr("value"))) : parse_vector(rad->attr("value"), 3, true);
                for (int k = 0; k < 3; ++k) a->radiance[k] = v[k];
                a->mesh = m->index;
                m->emitter = (int)emitters.size();
                emitters.push_back(a);
            }
            meshes.push_back(m);
        }
        auto add_all = [&](auto &arr, const char *name) {   // build_param_map, scene_loader.cpp:187-205
            for (size_t i = 0; i < arr.size(); ++i) {
                param_map[std::string(name) + "[" + std::to_string(i) + "]"] = arr[i];
                if (!arr[i]->id.empty()) {
                    const std::string key = std::string(name) + "[id=" + arr[i]->id + "]";
                    if (param_map.count(key)) throw std::runtime_error("Duplicate id: " + arr[i]->id);
                    param_map[key] = arr[i];
                }
            }
        };
        add_all(meshes, "Mesh"); add_all(emitters, "Emitter"); add_all(sensors, "Sensor");
        loaded = true;
        if (auto_configure) configure();
    }

    static int tex_slot(const RoughConductor &, int k) { return k; }
    void push_bitmap(int bsdf, int slot, Bitmap &b) {
        if (b.dirty) { check(pb_scene_set_bsdf_texture(ctx, bsdf, slot, b.data.data(), b.width, b.height)); b.dirty = false; }
        check(pb_grad_require(ctx, PB_PARAM_BSDF_TEXTURE, bsdf, slot, b.requires_grad ? 1 : 0));
    }

    void configure() {   // Scene::configure (scene.cpp:56-278): mirror the (possibly edited) objects into the context, then configure it
        if (!loaded) throw std::runtime_error("Scene not loaded yet!");
        if (!ctx) throw std::runtime_error("psdr_b200 needs a CUDA device (no CPU fallback): this Scene was created without a context");
        check(pb_scene_set_options(ctx, opts.width, opts.height, opts.spp, opts.sppe, opts.sppse));
        if (!uploaded) {
            for (auto &s : sensors) check_id(pb_scene_add_sensor(ctx, s->fov_x, s->near_clip, s->far_clip, s->to_world.m));
            for (auto &b : bsdfs) check_id(pb_scene_add_bsdf(ctx, dynamic_cast<Diffuse *>(b.get()) ? PB_BSDF_DIFFUSE : PB_BSDF_ROUGHCONDUCTOR));
            if (envmap) check_id(pb_scene_add_envmap(ctx, envmap->radiance.width, envmap->radiance.height, envmap->radiance.data.data(), envmap->scale, envmap->to_world_raw.m));
            if (envmap) { envmap->radiance.dirty = false; envmap->scale_dirty = false; }
            for (auto &m : meshes) {
                const int flags = (m->use_face_normals ? PB_MESH_FACE_NORMALS : 0) | (m->enable_edges ? PB_MESH_ENABLE_EDGES : 0);
                check_id(pb_scene_add_mesh(ctx, m->nv(), m->nf(), m->verts.data(), m->faces.data(), (int)m->uvs.size() / 2, m->uvs.empty() ? nullptr : m->uvs.data(),
                                           m->uv_faces.empty() ? nullptr : m->uv_faces.data(), flags, m->bsdf, m->to_world_raw.m));
                if (m->emitter >= 0) check_id(pb_scene_add_area_emitter(ctx, m->index, std::static_pointer_cast<AreaLight>(emitters[m->emitter])->radiance));
                m->verts_dirty = false;
                m->uploaded = true;
            }
            uploaded = true;
        }
        for (auto &s : sensors) if (s->dirty) { check(pb_scene_set_sensor_transform(ctx, (int)(&s - &sensors[0]), s->to_world.m)); s->dirty = false; }
        for (auto &s : sensors) check(pb_grad_require(ctx, PB_PARAM_SENSOR_TRANSFORM, (int)(&s - &sensors[0]), 0, s->requires_grad ? 1 : 0));
        for (auto &b : bsdfs) {
            if (auto *d = dynamic_cast<Diffuse *>(b.get())) push_bitmap(b->index, PB_TEX_REFLECTANCE, d->reflectance);
            else if (auto *r = dynamic_cast<RoughConductor *>(b.get())) {
                push_bitmap(b->index, PB_TEX_ALPHA_U, r->alpha_u); push_bitmap(b->index, PB_TEX_ALPHA_V, r->alpha_v);
                push_bitmap(b->index, PB_TEX_ETA, r->eta); push_bitmap(b->index, PB_TEX_K, r->k);
                push_bitmap(b->index, PB_TEX_SPECULAR_REFLECTANCE, r->specular_reflectance);
            }
        }
        for (auto &m : meshes) {
            if (m->verts_dirty) { check(pb_scene_set_mesh_vertices(ctx, m->index, m->verts.data())); m->verts_dirty = false; }
            if (m->transform_dirty) {
                check(pb_scene_set_mesh_transform(ctx, m->index, m->to_world_left.m, 1));
                check(pb_scene_set_mesh_transform(ctx, m->index, m->to_world_right.m, 0));
                m->transform_dirty = false;
            }
            check(pb_grad_require(ctx, PB_PARAM_MESH_VERTICES, m->index, 0, m->requires_grad ? 1 : 0));
            if (m->uvs_dirty && !m->uvs.empty()) { check(pb_scene_set_mesh_uvs(ctx, m->index, m->uvs.data())); m->uvs_dirty = false; }
            if (!m->uvs.empty()) check(pb_grad_require(ctx, PB_PARAM_MESH_UV, m->index, 0, m->uv_requires_grad ? 1 : 0));
        }
        if (envmap && envmap->transform_dirty) { check(pb_scene_set_envmap_transform(ctx, envmap->to_world_left.m)); envmap->transform_dirty = false; }
        if (envmap) {
            if (envmap->radiance.dirty || envmap->scale_dirty) {
                check(pb_scene_set_envmap_radiance(ctx, envmap->radiance.dirty ? envmap->radiance.data.data() : nullptr, envmap->scale));
                envmap->radiance.dirty = false; envmap->scale_dirty = false;
            }
            check(pb_grad_require(ctx, PB_PARAM_ENVMAP_RADIANCE, 0, 0, envmap->radiance.requires_grad ? 1 : 0));
            check(pb_grad_require(ctx, PB_PARAM_ENVMAP_SCALE, 0, 0, envmap->scale_requires_grad ? 1 : 0));
            check(pb_grad_require(ctx, PB_PARAM_ENVMAP_TRANSFORM, 0, 0, envmap->transform_requires_grad ? 1 : 0));
        }
        check(pb_scene_configure(ctx));
        configured = true;
    }

    // gradient vector layout as (param_map key, field, offset, count)
    py::list grad_layout() const {
        py::list out;
        const int n = pb_grad_num_segments(ctx);
        static const char *slots[] = {"reflectance", "alpha_u", "alpha_v", "eta", "k", "specular_reflectance"};
        for (int i = 0; i < n; ++i) {
            int kind, id, slot; int64_t off, cnt;
            check(pb_grad_segment(ctx, i, &kind, &id, &slot, &off, &cnt));
            if (kind == PB_PARAM_BSDF_TEXTURE) out.append(py::make_tuple("BSDF[" + std::to_string(id) + "]", std::string(slots[slot]), off, cnt));
            else if (kind == PB_PARAM_MESH_UV) out.append(py::make_tuple("Mesh[" + std::to_string(id) + "]", std::string("vertex_uv"), off, cnt));
            else if (kind == PB_PARAM_SENSOR_TRANSFORM) out.append(py::make_tuple("Sensor[" + std::to_string(id) + "]", std::string("to_world"), off, cnt));
            else if (kind == PB_PARAM_ENVMAP_TRANSFORM) out.append(py::make_tuple("Emitter[" + std::to_string(id) + "]", std::string("to_world_left"), off, cnt));
            else if (kind == PB_PARAM_ENVMAP_RADIANCE || kind == PB_PARAM_ENVMAP_SCALE)
                out.append(py::make_tuple("Emitter[" + std::to_string(id) + "]", std::string(kind == PB_PARAM_ENVMAP_RADIANCE ? "radiance" : "scale"), off, cnt));
            else out.append(py::make_tuple("Mesh[" + std::to_string(id) + "]", std::string("vertex_positions"), off, cnt));
        }
        return out;
    }
    std::string to_string() const {
        std::ostringstream oss;
        oss << "Scene[\n  # Sensors\n";
        for (auto &s : sensors) oss << "  " << s->to_string() << "\n";
        oss << "\n  # BSDFs\n";
        for (auto &b : bsdfs) oss << "  " << b->to_string() << "\n";
        oss << "\n  # Meshes\n";
        for (auto &m : meshes) oss << "  " << m->to_string() << "\n";
        oss << "]";
        return oss.str();
    }
};

struct Integrator {
    pb_integrator desc{PB_INTEG_DIRECT, 1, 1, 0, 0, 1, 0};
    virtual ~Integrator() = default;
    void require_ready(const Scene &s) const { if (!s.configured) throw std::runtime_error("Input scene must be configured!"); }
    // images are written into caller-provided device memory (W*H*3 floats); psdr_cuda/__init__.py allocates torch tensors
    void render_c(Scene &s, int sensor, uintptr_t d_image) const { require_ready(s); s.check(pb_render_c(s.ctx, &desc, sensor, reinterpret_cast<float *>(d_image))); }
    void render_d(Scene &s, int sensor, uintptr_t d_image) const { require_ready(s); s.check(pb_render_d(s.ctx, &desc, sensor, reinterpret_cast<float *>(d_image))); }
    void render_d_vjp(Scene &s, int sensor, uintptr_t d_dLdI, uintptr_t d_grad) const {
        require_ready(s);
        s.check(pb_render_d_vjp(s.ctx, &desc, sensor, reinterpret_cast<const float *>(d_dLdI), reinterpret_cast<float *>(d_grad)));
    }
    void render_d_jvp(Scene &s, int sensor, uintptr_t d_tangent, uintptr_t d_dimage) const {
        require_ready(s);
        s.check(pb_render_d_jvp(s.ctx, &desc, sensor, reinterpret_cast<const float *>(d_tangent), reinterpret_cast<float *>(d_dimage)));
    }
    farray render_c_numpy(Scene &s, int sensor) const {
        require_ready(s);
        farray img(std::vector<py::ssize_t>{(py::ssize_t)s.opts.width * s.opts.height, 3});
        s.check(pb_render_c_host(s.ctx, &desc, sensor, img.mutable_data()));
        return img;
    }
    void preprocess_secondary_edges(Scene &s, int sensor, const iarray &reso, int nrounds) {
        if (reso.size() != 4) throw std::runtime_error("resolution must have 4 entries");
        s.check(pb_preprocess_secondary_edges(s.ctx, sensor, reso.data(), nrounds));
        desc.use_guiding = 1;
    }
};
struct DirectIntegrator : Integrator {
    DirectIntegrator(int b, int l) {
        if (!(b >= 0 && l >= 0 && b + l > 0)) throw std::runtime_error("Invalid DirectIntegrator sample counts");
        desc.kind = PB_INTEG_DIRECT; desc.bsdf_samples = b; desc.light_samples = l;
    }
};
struct PathIntegrator : Integrator {
    explicit PathIntegrator(int depth) { if (depth < 1) throw std::runtime_error("PathIntegrator needs max_depth >= 1"); desc.kind = PB_INTEG_PATH; desc.max_depth = depth; }
};
struct FieldExtractionIntegrator : Integrator {
    explicit FieldExtractionIntegrator(const std::string &field) {
        static const std::map<std::string, int> f = {{"silhouette", PB_FIELD_SILHOUETTE}, {"position", PB_FIELD_POSITION}, {"depth", PB_FIELD_DEPTH},
                                                     {"geoNormal", PB_FIELD_GEONORMAL}, {"shNormal", PB_FIELD_SHNORMAL}, {"uv", PB_FIELD_UV}};
        auto it = f.find(field);
        if (it == f.end()) throw std::runtime_error("Unknown field: " + field);
        desc.kind = PB_INTEG_FIELD; desc.field = it->second;
    }
};

}  // namespace

PYBIND11_MODULE(_psdr_host, m) {
    m.doc() = "psdr-b200 host layer (pybind11 over the C ABI of include/psdr_b200.h)";
    m.def("abi_version", &pb_version);

    py::class_<Object, std::shared_ptr<Object>>(m, "Object")
        .def("type_name", &Object::type_name)
        .def_readonly("id", &Object::id)
        .def("__repr__", &Object::to_string);

    py::class_<RenderOption>(m, "RenderOption")
        .def(py::init<>())   // keyword names as in src/psdr.cpp:55-57
        .def(py::init<int, int, int>(), py::arg("width"), py::arg("height"), py::arg("spp/sppe"))
        .def(py::init<int, int, int, int>(), py::arg("width"), py::arg("height"), py::arg("spp"), py::arg("sppe"))
        .def(py::init<int, int, int, int, int>(), py::arg("width"), py::arg("height"), py::arg("spp"), py::arg("sppe"), py::arg("sppse"))
        .def_readwrite("width", &RenderOption::width).def_readwrite("height", &RenderOption::height).def_readwrite("spp", &RenderOption::spp)
        .def_readwrite("sppe", &RenderOption::sppe).def_readwrite("sppse", &RenderOption::sppse).def_readwrite("log_level", &RenderOption::log_level)
        .def("__repr__", [](const RenderOption &o) {   // src/psdr.cpp:63-71
            std::ostringstream s;
            s << "[width: " << o.width << ", height: " << o.height << ", spp: " << o.spp << ", sppe: " << o.sppe << ", sppse: " << o.sppse << ", log_level: " << o.log_level << "]";
            return s.str();
        });

    auto bitmap = [&](const char *name) {
        return py::class_<Bitmap>(m, name)
            .def(py::init([](int channels) { if (channels != 1 && channels != 3) throw std::runtime_error("Bitmap: 1 or 3 channels"); return Bitmap(channels, 0.f); }), py::arg("channels"))
            .def_property("data", &Bitmap::get_data, &Bitmap::set_data)
            .def_property("resolution", [](const Bitmap &b) { return py::make_tuple(b.width, b.height); },
                          [](Bitmap &b, std::pair<int, int> r) { b.width = r.first; b.height = r.second; b.data.assign((size_t)b.width * b.height * b.channels, 0.f); b.dirty = true; })
            .def_readwrite("requires_grad", &Bitmap::requires_grad)
            .def_readonly("channels", &Bitmap::channels);
    };
    bitmap("BitmapD");

    py::class_<BSDF, Object, std::shared_ptr<BSDF>>(m, "BSDF").def_readonly("index", &BSDF::index)
        .def("anisotropic", [](const BSDF &) { return false; });   // bsdf.h:33; every BSDF the loader / Python can create is isotropic (roughconductor.h:11-18)
    py::class_<Diffuse, BSDF, std::shared_ptr<Diffuse>>(m, "DiffuseBSDF")
        .def_property("reflectance", py::cpp_function([](Diffuse &d) -> Bitmap & { return d.reflectance; }, py::return_value_policy::reference_internal), [](Diffuse &d, const Bitmap &b) { d.reflectance.assign(b); });
    py::class_<RoughConductor, BSDF, std::shared_ptr<RoughConductor>>(m, "RoughConductorBSDF")
        .def_property("alpha_u", py::cpp_function([](RoughConductor &d) -> Bitmap & { return d.alpha_u; }, py::return_value_policy::reference_internal), [](RoughConductor &d, const Bitmap &b) { d.alpha_u.assign(b); })
        .def_property("alpha_v", py::cpp_function([](RoughConductor &d) -> Bitmap & { return d.alpha_v; }, py::return_value_policy::reference_internal), [](RoughConductor &d, const Bitmap &b) { d.alpha_v.assign(b); })
        .def_property("eta", py::cpp_function([](RoughConductor &d) -> Bitmap & { return d.eta; }, py::return_value_policy::reference_internal), [](RoughConductor &d, const Bitmap &b) { d.eta.assign(b); })
        .def_property("k", py::cpp_function([](RoughConductor &d) -> Bitmap & { return d.k; }, py::return_value_policy::reference_internal), [](RoughConductor &d, const Bitmap &b) { d.k.assign(b); })
        .def_property("specular_reflectance", py::cpp_function([](RoughConductor &d) -> Bitmap & { return d.specular_reflectance; }, py::return_value_policy::reference_internal), [](RoughConductor &d, const Bitmap &b) { d.specular_reflectance.assign(b); });

    py::class_<Sensor, Object, std::shared_ptr<Sensor>>(m, "PerspectiveCamera")
        .def(py::init([](float fov_x, float near_clip, float far_clip) { auto s = std::make_shared<Sensor>(); s->fov_x = fov_x; s->near_clip = near_clip; s->far_clip = far_clip; return s; }),
             py::arg("fov_x"), py::arg("near"), py::arg("far"))   // src/psdr.cpp:223
        .def_property("to_world", [](const Sensor &s) { return mat_to_numpy(s.to_world); }, [](Sensor &s, const farray &a) { s.to_world = mat_from_numpy(a); s.dirty = true; })
        .def_readwrite("requires_grad", &Sensor::requires_grad)
        .def_readonly("fov_x", &Sensor::fov_x).def_readonly("near_clip", &Sensor::near_clip).def_readonly("far_clip", &Sensor::far_clip);

    py::class_<Emitter, Object, std::shared_ptr<Emitter>>(m, "Emitter");
    py::class_<AreaLight, Emitter, std::shared_ptr<AreaLight>>(m, "AreaLight")
        .def(py::init([](const std::array<float, 3> &radiance, const Mesh *mesh) {   // src/psdr.cpp:231
                 auto a = std::make_shared<AreaLight>(); for (int k = 0; k < 3; ++k) a->radiance[k] = radiance[k]; a->mesh = mesh ? mesh->index : -1; return a; }),
             py::arg("radiance"), py::arg("mesh"))
        .def_property_readonly("radiance", [](const AreaLight &a) { return py::make_tuple(a.radiance[0], a.radiance[1], a.radiance[2]); });
    py::class_<EnvironmentMap, Emitter, std::shared_ptr<EnvironmentMap>>(m, "EnvironmentMap")
        .def(py::init([]() { return std::make_shared<EnvironmentMap>(); }))   // the Python layer adds the (file name) form of src/psdr.cpp:234
        .def_property("radiance", py::cpp_function([](EnvironmentMap &e) -> Bitmap & { return e.radiance; }, py::return_value_policy::reference_internal), [](EnvironmentMap &e, const Bitmap &b) { e.radiance.assign(b); })
        .def_property("scale", [](const EnvironmentMap &e) { return e.scale; }, [](EnvironmentMap &e, float v) { e.scale = v; e.scale_dirty = true; })
        .def_readwrite("scale_requires_grad", &EnvironmentMap::scale_requires_grad)
        .def_readwrite("transform_requires_grad", &EnvironmentMap::transform_requires_grad)
        .def_property_readonly("to_world_left", [](const EnvironmentMap &e) { return mat_to_numpy(e.to_world_left); })
        .def_property_readonly("to_world", [](const EnvironmentMap &e) { return mat_to_numpy(e.to_world_left * e.to_world_raw); })
        .def("set_transform", [](EnvironmentMap &e, const farray &a) { e.to_world_left = mat_from_numpy(a); e.transform_dirty = true; });

    py::class_<Mesh, Object, std::shared_ptr<Mesh>>(m, "Mesh")
        .def(py::init<>())                                                      // src/psdr.cpp:242-243
        .def("load", [](Mesh &x, const std::string &path, bool) { x.load(path); }, py::arg("filename"), py::arg("verbose") = false)
        .def_property("face_uv_indices", [](const Mesh &x) {
            iarray a(std::vector<py::ssize_t>{(py::ssize_t)(x.uv_faces.size() / 3), 3});
            if (!x.uv_faces.empty()) std::memcpy(a.mutable_data(), x.uv_faces.data(), x.uv_faces.size() * sizeof(int));
            return a; }, &Mesh::set_uv_faces)
        .def_property_readonly("num_vertices", &Mesh::nv)
        .def_property_readonly("num_faces", &Mesh::nf)
        .def_property("vertex_positions", &Mesh::get_vertices, &Mesh::set_vertices)
        .def_property("vertex_uv", &Mesh::get_uvs, &Mesh::set_uvs)
        .def_property("face_indices", &Mesh::get_faces, &Mesh::set_faces)
        .def_property_readonly("to_world_raw", [](const Mesh &x) { return mat_to_numpy(x.to_world_raw); })
        .def_property_readonly("to_world_left", [](const Mesh &x) { return mat_to_numpy(x.to_world_left); })
        .def_property_readonly("to_world_right", [](const Mesh &x) { return mat_to_numpy(x.to_world_right); })
        .def_readonly("index", &Mesh::index)
        .def_readonly("bsdf_index", &Mesh::bsdf)
        .def_readonly("emitter_index", &Mesh::emitter)
        .def_property_readonly("has_uv", [](const Mesh &x) { return !x.uvs.empty(); })
        .def_readwrite("enable_edges", &Mesh::enable_edges)
        .def_readwrite("use_face_normals", &Mesh::use_face_normals)
        .def_readwrite("requires_grad", &Mesh::requires_grad)
        .def_readwrite("uv_requires_grad", &Mesh::uv_requires_grad)
        .def_property_readonly("to_world", [](const Mesh &x) { return mat_to_numpy(x.to_world_left * x.to_world_raw * x.to_world_right); })
        .def("set_transform", [](Mesh &x, const farray &a, bool set_left) { (set_left ? x.to_world_left : x.to_world_right) = mat_from_numpy(a); x.transform_dirty = true; },
             py::arg("mat"), py::arg("set_left") = true)   // mesh.h:19-26
        .def("append_transform", [](Mesh &x, const farray &a, bool append_left) {   // mesh.h:28-35
                 if (append_left) x.to_world_left = mat_from_numpy(a) * x.to_world_left; else x.to_world_right = x.to_world_right * mat_from_numpy(a);
                 x.transform_dirty = true;
             }, py::arg("mat"), py::arg("append_left") = true)
        .def("dump", &Mesh::dump);

    py::class_<Scene>(m, "Scene")
        .def(py::init<int>(), py::arg("device") = 0)
        .def("load_file", &Scene::load_file, py::arg("file_name"), py::arg("auto_configure") = true)
        .def("load_string", &Scene::load_string, py::arg("scene_xml"), py::arg("auto_configure") = true)
        .def("configure", &Scene::configure)
        .def_readwrite("opts", &Scene::opts)
        .def_property_readonly("num_sensors", [](const Scene &s) { return (int)s.sensors.size(); })
        .def_property_readonly("num_meshes", [](const Scene &s) { return s.configured ? pb_scene_num_meshes(s.ctx) : (int)s.meshes.size(); })
        .def_property_readonly("param_map", [](Scene &s) { py::dict d; for (auto &kv : s.param_map) d[py::str(kv.first)] = py::cast(kv.second); return d; })
        .def("grad_layout", &Scene::grad_layout)
        .def("grad_size", [](const Scene &s) { return (int64_t)pb_grad_size(s.ctx); })
        .def("set_shard", [](Scene &s, int rank, int world) { s.check(pb_ctx_set_shard(s.ctx, rank, world)); s.shard_world = world; })
        .def("set_shard_mode", [](Scene &s, const std::string &mode, int tile_rows) {
                 if (mode != "samples" && mode != "pixels") throw std::runtime_error("shard mode must be \"samples\" or \"pixels\"");
                 s.check(pb_ctx_set_shard_mode(s.ctx, mode == "pixels" ? PB_SHARD_PIXELS : PB_SHARD_SAMPLES, tile_rows));
             }, py::arg("mode"), py::arg("tile_rows") = 0)
        // several GPUs: the library owns an NCCL communicator and enqueues the film / gradient collectives on the scene's stream
        .def_static("dist_available", []() { return pb_dist_available() != 0; })
        .def("dist_unique_id", [](Scene &s) { char id[128]; s.check(pb_dist_unique_id(s.ctx, id)); return py::bytes(id, 128); })
        .def("dist_init", [](Scene &s, const std::string &id, int rank, int world) {
                 if (id.size() != 128) throw std::runtime_error("the NCCL unique id has 128 bytes");
                 s.check(pb_dist_init(s.ctx, id.data(), rank, world)); s.shard_world = world;
             }, py::arg("unique_id"), py::arg("rank"), py::arg("world"))
        .def("dist_finalize", [](Scene &s) { s.check(pb_dist_finalize(s.ctx)); })
        .def_property_readonly("shard_world", [](const Scene &s) { return s.shard_world; })
        .def("_allreduce_image", [](Scene &s, uintptr_t d_image) { s.check(pb_allreduce_image(s.ctx, reinterpret_cast<float *>(d_image))); })
        .def("_allreduce_grads", [](Scene &s, uintptr_t d_grad, int64_t n) { s.check(pb_allreduce_grads(s.ctx, reinterpret_cast<float *>(d_grad), n)); })
        .def("_render_d_state", [](Scene &s) { uint64_t st[4]; s.check(pb_render_d_get_state(s.ctx, st)); return py::make_tuple(st[0], st[1], st[2], st[3]); })
        .def("_render_d_set_state", [](Scene &s, const std::vector<uint64_t> &st) {
                 if (st.size() != 4) throw std::runtime_error("a renderD state has 4 entries");
                 s.check(pb_render_d_set_state(s.ctx, st.data()));
             })
        .def("_sample_boundary_segment_direct", [](Scene &s, int64_t n, uintptr_t d_sample3, uintptr_t d_out) {
                 if (!s.configured) throw std::runtime_error("Scene needs to be configured!");
                 s.check(pb_sample_boundary_segment_direct(s.ctx, n, reinterpret_cast<const float *>(d_sample3), reinterpret_cast<float *>(d_out)));
             })
        // parameters that live on the GPU (torch leaves): device-to-device updates, no host round trip (pb_scene_set_*_device)
        .def("_set_vertices_device", [](Scene &s, int mesh, uintptr_t d_verts) {
                 if (!s.uploaded) throw std::runtime_error("the scene has not been configured yet");
                 s.check(pb_scene_set_mesh_vertices_device(s.ctx, mesh, reinterpret_cast<const float *>(d_verts)));
                 s.meshes.at(mesh)->verts_dirty = false;
             })
        .def("_get_vertices_device", [](Scene &s, int mesh) {
                 auto &m = s.meshes.at(mesh);
                 farray a(std::vector<py::ssize_t>{(py::ssize_t)m->nv(), 3});
                 s.check(pb_scene_get_mesh_vertices(s.ctx, mesh, a.mutable_data()));
                 std::memcpy(m->verts.data(), a.data(), m->verts.size() * sizeof(float));   // the host mirror follows, without marking it edited
                 return a;
             })
        .def("_set_texture_device", [](Scene &s, int bsdf, int slot, uintptr_t d_data) {
                 if (!s.uploaded) throw std::runtime_error("the scene has not been configured yet");
                 s.check(pb_scene_set_bsdf_texture_device(s.ctx, bsdf, slot, reinterpret_cast<const float *>(d_data)));
             })
        .def("set_edge_importance", [](Scene &s, const std::string &mode) {
                 if (mode != "length" && mode != "dihedral") throw std::runtime_error("edge importance must be \"length\" or \"dihedral\"");
                 s.check(pb_scene_set_edge_importance(s.ctx, mode == "dihedral" ? 1 : 0));
             })
        .def_property_readonly("uploaded", [](const Scene &s) { return s.uploaded; })
        .def("stats_collectives", [](const Scene &s) { return (int64_t)pb_stats_collectives(s.ctx); })
        .def("set_stream", [](Scene &s, uintptr_t stream) { s.check(pb_ctx_set_stream(s.ctx, reinterpret_cast<void *>(stream))); })
        .def("stats_launches", [](const Scene &s) { return (int64_t)pb_stats_launches(s.ctx); })
        .def("__repr__", &Scene::to_string);

    py::class_<Integrator>(m, "Integrator")
        .def("_render_c", &Integrator::render_c)
        .def("_render_d", &Integrator::render_d)
        .def("_render_d_vjp", &Integrator::render_d_vjp)
        .def("_render_d_jvp", &Integrator::render_d_jvp)
        .def("renderC_numpy", &Integrator::render_c_numpy, py::arg("scene"), py::arg("sensor_id") = 0)
        .def("preprocess_secondary_edges", &Integrator::preprocess_secondary_edges, py::arg("scene"), py::arg("sensor_id"), py::arg("resolution"), py::arg("nrounds") = 1)
        .def_property("hide_emitters", [](const Integrator &i) { return i.desc.hide_emitters != 0; }, [](Integrator &i, bool v) { i.desc.hide_emitters = v ? 1 : 0; });
    py::class_<DirectIntegrator, Integrator>(m, "DirectIntegrator").def(py::init<int, int>(), py::arg("bsdf_samples") = 1, py::arg("light_samples") = 1);
    py::class_<PathIntegrator, Integrator>(m, "PathIntegrator").def(py::init<int>(), py::arg("max_depth") = 5);
    py::class_<FieldExtractionIntegrator, Integrator>(m, "FieldExtractionIntegrator").def(py::init<std::string>(), py::arg("field"));
}
