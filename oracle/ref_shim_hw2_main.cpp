// ref_shim_hw2_main.cpp — the reference's whole bvh_viz program (HW2/HW2/GPUandCPU/src/main.cu, CPU
// build) reachable in-process: main() is renamed by the preprocessor and the file is #included where
// it lies, which also exposes its file-static helpers (applyObjectTransform, main.cu:75-96).
// TEST INFRASTRUCTURE ONLY; output oracle/_ref/libref_hw2_main.so (git-ignored).  Used by
// tools/make_golden.py to produce end-to-end fixtures (scene JSON + OBJ in, 8-bit image out).
#include <cstring>
#define main ref_bvh_viz_main
#include "main.cu"
#undef main

extern "C" {

// Runs the unmodified program; it writes render.png into the current directory.
int ref_hw2_main(int argc, char** argv) { return ref_bvh_viz_main(argc, argv); }

// applyObjectTransform on raw arrays (nrm may be NULL).
void ref_hw2_transform(float* pos, float* nrm, uint64_t nv, const float* position, const float* rotation, const float* scale) {
    Mesh m;
    m.positions.assign((Vec3*)pos, (Vec3*)pos + nv);
    if (nrm) m.normals.assign((Vec3*)nrm, (Vec3*)nrm + nv);
    SceneObject o;
    o.position = make_vec3(position[0], position[1], position[2]);
    o.rotation = make_vec3(rotation[0], rotation[1], rotation[2]);
    o.scale = make_vec3(scale[0], scale[1], scale[2]);
    applyObjectTransform(m, o);
    std::memcpy(pos, m.positions.data(), nv * sizeof(Vec3));
    if (nrm) std::memcpy(nrm, m.normals.data(), nv * sizeof(Vec3));
}

} // extern "C"
