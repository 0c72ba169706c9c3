// TEST INFRASTRUCTURE (oracle/): a thin C shim around the reference's OWN vendored ingest code, compiled from the sources where they
// lie under /root/reference (include/tiny_obj_loader/tiny_obj_loader.h, include/psdr/core/tinyexr.h, src/core/miniz.cpp) by
// oracle/build_ref.sh into oracle/_ref/libref_ingest.so. Nothing of the reference is copied into this repository.
//
// The rendering path of psdr-cuda cannot be built here (Enoki + OptiX); its file ingest can: these two functions call
// tinyobj::LoadObj and LoadEXR exactly the way Mesh::load (src/shape/mesh.cpp:62-141) and BitmapLoader::load_openexr_rgba
// (src/core/bitmap_loader.cpp:13-53) do, so the oracle's and the product's loaders can be pinned to the reference's own parsers.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader/tiny_obj_loader.h>

#define TINYEXR_USE_MINIZ 0
#include <psdr/core/miniz.h>
#define TINYEXR_IMPLEMENTATION
#include <psdr/core/tinyexr.h>

namespace {
struct ObjData { std::vector<float> verts, uvs; std::vector<int> faces, uv_faces; };
ObjData g_obj;
std::vector<float> g_exr;
}

extern "C" {

// mesh.cpp:62-141: vertices / texcoords as tinyobj returns them, one (vertex_index, texcoord_index) triple per face corner
int ref_load_obj(const char *path, int *nv, int *nuv, int *nf) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, err;
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &err, path)) return -1;
    g_obj = ObjData();
    g_obj.verts.assign(attrib.vertices.begin(), attrib.vertices.end());
    g_obj.uvs.assign(attrib.texcoords.begin(), attrib.texcoords.end());
    const bool has_uv = !attrib.texcoords.empty();
    for (size_t s = 0; s < shapes.size(); ++s)
        for (size_t f = 0; f < shapes[s].mesh.num_face_vertices.size(); ++f) {
            if (shapes[s].mesh.num_face_vertices[f] != 3) return -2;   // mesh.cpp:121 asserts triangles (tinyobj triangulates by default)
            for (int i = 0; i < 3; ++i) {
                const tinyobj::index_t idx = shapes[s].mesh.indices[3 * f + i];
                g_obj.faces.push_back(idx.vertex_index);
                if (has_uv) g_obj.uv_faces.push_back(idx.texcoord_index);
            }
        }
    *nv = (int)g_obj.verts.size() / 3; *nuv = (int)g_obj.uvs.size() / 2; *nf = (int)g_obj.faces.size() / 3;
    return 0;
}
void ref_get_obj(float *verts, float *uvs, int *faces, int *uv_faces) {
    if (verts) std::memcpy(verts, g_obj.verts.data(), g_obj.verts.size() * sizeof(float));
    if (uvs) std::memcpy(uvs, g_obj.uvs.data(), g_obj.uvs.size() * sizeof(float));
    if (faces) std::memcpy(faces, g_obj.faces.data(), g_obj.faces.size() * sizeof(int));
    if (uv_faces) std::memcpy(uv_faces, g_obj.uv_faces.data(), g_obj.uv_faces.size() * sizeof(int));
}

// bitmap_loader.cpp:13-53: RGBA float image, row-major from the top
int ref_load_exr(const char *path, int *w, int *h) {
    float *out = nullptr;
    const char *err = nullptr;
    if (LoadEXR(&out, w, h, path, &err) != TINYEXR_SUCCESS) { if (err) FreeEXRErrorMessage(err); return -1; }
    g_exr.assign(out, out + (size_t)(*w) * (*h) * 4);
    free(out);
    return 0;
}
void ref_get_exr(float *rgba) { std::memcpy(rgba, g_exr.data(), g_exr.size() * sizeof(float)); }

}  // extern "C"
