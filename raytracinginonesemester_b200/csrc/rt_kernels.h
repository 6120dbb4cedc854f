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
// Writes the first position i with idx[i] >= num_vertices into *bad (device, preset to ~0 by the caller).
cudaError_t rt_validate_indices(const uint32_t* idx, size_t n, uint32_t num_vertices, unsigned long long* bad, cudaStream_t stream);
// Bakes an object's transform into its vertex range in place (device arrays), before the build.
struct BakeXform;
cudaError_t rt_bake_transform(float* pos, float* nrm, size_t first, size_t count, const BakeXform& T, cudaStream_t stream);
// 8-wide view of a finished BVH2 for the frustum traversal: one WideNode per BVH2 node at depth 0, 3, 6, ... (compact array,
// allocated here with cudaMalloc; *count_out nodes; references between wide nodes are indices into it; the root is wide[0]).
cudaError_t rt_build_wide(const BvhNode* nodes, uint32_t num_nodes, int phase, WideNode** wide_out, uint32_t* count_out, cudaStream_t stream);
// Pack triangles in input order without a BVH (brute-force-only scenes).
cudaError_t rt_pack_triangles(const BuildParams& bp, TriBlock* geom, TriBlock* shade, cudaStream_t stream);

// rt_trace.cu --------------------------------------------------------------------------
cudaError_t rt_launch_render(const FrameParams& fp, int kernel_variant, cudaStream_t stream, int* launches);
// Which frames run on the persistent kernel, which publishes band-completion flags and runs the multi-GPU handshake itself
// (FrameParams.queue / flags / ready_* / wait_ranks); other kernels need the flag kernels below around them.  `banded`: the
// caller wants bands published while the frame renders AND cannot launch them one by one (rt_render_into of a multi-GPU context
// into rank 0's own buffer: the banded peer-store gather).  RT_VARIANT_DEFAULT picks the persistent kernel for those frames and the
// block-per-tile launch of the same traversal otherwise (single GPU and shared-host-image frames: one launch per band) — measured,
// B200, C4: device time 1.847 vs 1.954 ms on one GPU, 0.336 vs 0.374 ms per
// frame on eight; end to end on one GPU (one launch per band + event-chained copies vs one persistent launch + band flags) 1.96
// vs 2.03 ms: the hardware's block scheduler does the same job without a barrier per tile (11 % of the persistent kernel's stall
// samples) and tolerates one more resident block per SM.  The host stores the answer in FrameParams.persist.
bool rt_render_is_persistent(const FrameParams& fp, int kernel_variant, bool banded);
#define RT_PERSIST_CTL_BYTES 128          // device bytes behind FrameParams.queue
// Scatter tile-packed planes of rank `src_rank` into the row-major image (rank 0, world > 1).
cudaError_t rt_launch_unpack(const FrameParams& fp, int src_rank, const float* rgb, const uint8_t* rgb8,
                             const int32_t* tri_id, const float* t, float* o_rgb, uint8_t* o_rgb8,
                             int32_t* o_tri_id, float* o_t, cudaStream_t stream);
// Completion flags of the peer-store gather (see k_flag_set / k_flag_wait in rt_trace.cu).
#define RT_PEER_FLAG_STRIDE 16            // one 64-byte line per flag
#define RT_PEER_MAX_RANKS 8
#define RT_PEER_MAX_CHUNKS RT_MAX_BANDS   // bands per rank whose completion is published (rt_params.h)
// flag block layout (uints, each flag on its own line): [0] ready (rank 0 -> everybody: the previous image has been read),
// [16 + r*RT_PEER_MAX_CHUNKS + j] band j of rank r done
#define RT_PEER_CHUNK_FLAG(r, j) ((size_t)(16 + (r) * RT_PEER_MAX_CHUNKS + (j)) * RT_PEER_FLAG_STRIDE)
// (only around kernels that do not publish their own completion: brute force, per-ray, the block-per-tile variants)
cudaError_t rt_launch_flag_set(unsigned* flags, int stride, int n, unsigned seq, cudaStream_t stream);
// waits for flags[a * rank_stride + b * stride] >= seq for a < nranks, b < nper
cudaError_t rt_launch_flag_wait(const unsigned* flags, int rank_stride, int nranks, int stride, int nper, unsigned seq,
                                unsigned long long timeout_ns, unsigned* err, cudaStream_t stream);
cudaError_t rt_launch_flag_unblock(const unsigned* err, unsigned* flags, int stride, int n, unsigned seq, cudaStream_t stream);
