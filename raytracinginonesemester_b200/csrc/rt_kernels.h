// rt_kernels.h — launch interface between the C-ABI host layer (rt_api.cu) and the
// device code (rt_build.cu, rt_trace.cu).
#pragma once

#include <cuda_runtime.h>
#include "rt_params.h"

// rt_build.cu --------------------------------------------------------------------------
// Device LBVH build.  On return (stream-ordered) nodes/geom/shade hold the flattened BVH;
// *num_nodes_out (host) is valid after the call (it synchronises once to size the node array).
struct BuildResult { uint32_t num_nodes; uint32_t num_leaves; float scene_min[3], scene_max[3]; };
cudaError_t rt_build_bvh(const BuildParams& bp, BvhNode** nodes_out, TriBlock* geom, TriBlock* shade,
                         BuildResult* res, cudaStream_t stream);
// Bakes an object's transform into its vertex range in place (device arrays), before the build.
struct BakeXform;
cudaError_t rt_bake_transform(float* pos, float* nrm, size_t first, size_t count, const BakeXform& T, cudaStream_t stream);
// 8-wide view of a finished BVH2 (one WideNode per BVH2 node), for the frustum traversal.
cudaError_t rt_build_wide(const BvhNode* nodes, uint32_t num_nodes, WideNode* wide, cudaStream_t stream);
// Pack triangles in input order without a BVH (brute-force-only scenes).
cudaError_t rt_pack_triangles(const BuildParams& bp, TriBlock* geom, TriBlock* shade, cudaStream_t stream);

// rt_trace.cu --------------------------------------------------------------------------
cudaError_t rt_launch_render(const FrameParams& fp, int kernel_variant, cudaStream_t stream, int* launches);
// Scatter tile-packed planes of rank `src_rank` into the row-major image (rank 0, world > 1).
cudaError_t rt_launch_unpack(const FrameParams& fp, int src_rank, const float* rgb, const uint8_t* rgb8,
                             const int32_t* tri_id, const float* t, float* o_rgb, uint8_t* o_rgb8,
                             int32_t* o_tri_id, float* o_t, cudaStream_t stream);
// Completion flags of the peer-store gather (see k_flag_set / k_flag_wait in rt_trace.cu).
#define RT_PEER_FLAG_STRIDE 16            // one 64-byte line per flag
#define RT_PEER_MAX_RANKS 8
#define RT_PEER_MAX_CHUNKS 16            // ownership chunks per rank that rt_render_into can pipeline
// flag block layout (uints, each flag on its own line): [0] ready, [r] frame done by rank r (r >= 1),
// [16 + r*RT_PEER_MAX_CHUNKS + j] chunk j of rank r done
#define RT_PEER_CHUNK_FLAG(r, j) ((size_t)(16 + (r) * RT_PEER_MAX_CHUNKS + (j)) * RT_PEER_FLAG_STRIDE)
cudaError_t rt_launch_flag_set(unsigned* flag, unsigned seq, cudaStream_t stream);
cudaError_t rt_launch_flag_wait(const unsigned* flags, int stride, int n, unsigned seq, unsigned long long timeout_ns,
                                unsigned* err, cudaStream_t stream);
cudaError_t rt_launch_flag_unblock(const unsigned* err, unsigned* flags, int stride, int n, unsigned seq, cudaStream_t stream);
