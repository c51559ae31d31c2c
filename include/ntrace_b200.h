/* ntrace_b200 — C ABI of the B200-native tracing / BVH-build path (libntrace_b200.so).
 *
 * This is the drop-in boundary for NTrace's tracing path.  The reference has no FFI; its
 * de-facto operator API is the host virtual interface FW::CudaVirtualTracer
 * (src/rt/cuda/CudaVirtualTracer.hpp:11-26) plus the kernel plug-in ABI of
 * src/rt/kernels/CudaTracerKernels.hpp:69-112.  Each entry point below names the reference
 * interface it replaces.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; nt_last_error() gives the message
 *     (reference: sticky setError/getError, src/framework/base/Defs.hpp:143-147);
 *   - no exceptions cross the boundary; all pointers are caller-owned and borrowed for the call;
 *   - pointers may be HOST or DEVICE addresses (the reference's FW::Buffer migrates lazily,
 *     src/framework/gpu/Buffer.hpp:107-113); the library detects which with
 *     cudaPointerGetAttributes; page-locked (pinned, device-mapped) host ray / result buffers are traversed in place over
 *     PCIe, pageable ones are staged through the library's own device buffers;
 *   - calls are synchronous: when a call returns its outputs are complete
 *     (reference: CudaKernel::launchTimed syncs, src/framework/gpu/CudaKernel.cpp:188-221).  The two opt-in exceptions
 *     are nt_set_deferred(1 / 2) (device buffers only) and nt_trace_batch_async / nt_trace_wait;
 *   - one device per process (reference: one CUDA context, CudaModule.hpp:92-97); multi-GPU runs
 *     use one process per GPU and replicate the BVH (nt_bvh_device_ptrs + NCCL broadcast in the host);
 *   - not re-entrant: one caller thread at a time (guarded by an internal mutex);
 *   - there is NO CPU fallback: without a CUDA device every call fails with an error.
 */
#ifndef NTRACE_B200_H
#define NTRACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* BVHLayout values — identical to src/rt/kernels/CudaTracerKernels.hpp:52-63 */
#define NT_LAYOUT_AOS_AOS  0
#define NT_LAYOUT_AOS_SOA  1
#define NT_LAYOUT_SOA_AOS  2
#define NT_LAYOUT_SOA_SOA  3
#define NT_LAYOUT_COMPACT  4
#define NT_LAYOUT_COMPACT2 5
#define NT_LAYOUT_CPU      6

/* builders accepted by nt_bvh_build — HLBVHBuilder.cpp:44-47 (LBVH when !hlbvh or hlbvhBits == 10) */
#define NT_BUILDER_LBVH  0
#define NT_BUILDER_HLBVH 1

/* ---- lifetime / device ------------------------------------------------------------------- */
/* CudaModule::staticInit (src/framework/gpu/CudaModule.cpp:311): select the device, create the stream. */
int nt_init(int device_ordinal);
void nt_shutdown(void);
/* setError/getError (src/framework/base/Defs.hpp:143-147). Thread-local, never NULL. */
const char* nt_last_error(void);
/* number of this library's kernel launches since nt_init (bench.py's gpu_launches). */
int64_t nt_launch_count(void);

/* Memory (reference: FW::Buffer over cuMemAlloc / cuMemAllocHost / cuMemcpy*, src/framework/gpu/Buffer.cpp:383-520):
 * the C++ host above this ABI (ntrace_b200/host_cpp) keeps its FW::Buffer mirrors through these calls and links no
 * CUDA library itself.  nt_memcpy takes any mix of host / device pointers and returns when the copy is complete. */
int nt_mem_alloc(size_t bytes, void** outDevicePtr);
int nt_mem_free(void* devicePtr);
int nt_mem_alloc_host(size_t bytes, void** outHostPtr);      /* page-locked, device-mapped */
int nt_mem_free_host(void* hostPtr);
int nt_memcpy(void* dst, const void* src, size_t bytes);
int nt_memset(void* devicePtr, int value, size_t bytes);
/* CUDA-event timing on the library's stream (reference: CudaKernel::launchTimed brackets launches with
 * cuEventRecord / cuEventElapsedTime, src/framework/gpu/CudaKernel.cpp:188-221).  slot in [0, 8). */
int nt_event_record(int slot);
/* waits for slotB, then returns the device time between the two records. */
int nt_event_elapsed(int slotA, int slotB, float* outSeconds);
/* Submission mode.  0 (default) = reference behaviour, every call returns with its outputs complete.
 * 1 = deferred: nt_trace_batch on DEVICE buffers only enqueues (outSeconds = 0) and nt_synchronize() waits;
 * calls with host buffers stay synchronous.
 * 2 = deferred and overlapped: as 1, and consecutive launches alternate between two kernel streams, so that the CTAs of
 * the next batch move in while the persistent CTAs of the current one run out of rays (the batches of a frame are
 * independent once the primary results exist; the reference traces them one after the other, Renderer.cpp:405-579).
 * A launch is ordered after everything queued before it, and after any launch still in flight that shares a ray or result buffer with
 * it (the library remembers the buffers of the last 8 launches), so tracing the same batch twice is safe, merely not overlapped.
 * nt_raygen_ao on device buffers is queued too in this mode (it returns at once) and waits only for the launches in flight that use
 * ITS buffers: a host loop that alternates two secondary RayBuffers (generate B while A is traced) keeps the GPU full without any
 * call of its own to order things.  Every other call of this API waits for all launches in flight before it enqueues or reads
 * anything.  Mode 2 also queues calls whose buffers are page-locked host memory (traversed in place over PCIe; results are in host
 * memory once nt_synchronize() returns); pageable host buffers stay synchronous in every mode. */
int nt_set_deferred(int mode);
int nt_synchronize(void);

/* ---- kernel selection: CudaBVHTracer::setKernel + queryConfig (CudaBVHTracer.cpp:52-84) --- */
/* Names: "b200_persistent_speculative_while_while" (default), "b200_speculative_while_while": the triangle test in IEEE
 * arithmetic, bit-identical to the reference's CPU tracer (CudaBVH::trace, Util.cpp:99-127).  A "_fastmath" suffix
 * ("b200_persistent_speculative_while_while[_compact2]_fastmath") and the reference's own kernel file names select the
 * arithmetic nvcc -use_fast_math gives the reference's GPU kernels (contracted FMAs, approximate reciprocal), bit-identical to
 * those kernels recompiled for sm_100a.  Reference names accepted as aliases: "fermi_speculative_while_while" (Compact),
 * "kepler_dynamic_fetch" (Compact2), "tesla_persistent_while_while" / "tesla_persistent_speculative_while_while" /
 * "tesla_persistent_packet" (AOS_AOS, as those files ship); "b200_persistent_speculative_while_while_{aos_aos,aos_soa,
 * soa_aos,soa_soa}" select the other basic layouts.  "b200_wide4" / "b200_wide4_fastmath" (Compact) and "b200_wide4_compact2"
 * traverse a 4-wide, 8-bit quantised node array that the library derives from the resident Compact / Compact2 BVH
 * (csrc/nt_wide.cu): same leaves, same triangle test, so closest-hit t/u/v are bit-identical per (ray, triangle) and ids differ
 * from the binary kernels only on exact-t ties; any-hit rays report A hit (possibly another triangle than the binary order finds).
 * "b200_auto" (/ "_compact2") picks per batch: any-hit batches go through the binary kernel (reference visiting order), closest-hit batches
 * through Wide4 -- except camera rays the library generated itself (nt_raygen_primary into a device buffer that is then traced in place):
 * rays with one origin stay together far down the tree, where the cheaper binary node step wins.  "b200_mr*" / "b200_wide4_mr*" (both rays of a lane in shared memory, phase vote per step) and "b200_sw*" /
 * "b200_wide4_sw*" (one ray in registers, one parked in shared memory, exchanged at the phase boundaries) are the two-rays-per-lane
 * experiments: bit-identical to their one-ray twins, slower.
 * Unknown names fail. */
int nt_set_kernel(const char* name);
/* CudaBVHTracer::getDesiredBVHLayout -> BVHLayout of the selected kernel. */
int nt_desired_layout(void);
/* KernelConfig {bvhLayout, blockWidth, blockHeight, usePersistentThreads} (CudaTracerKernels.hpp:69-75). */
int nt_kernel_config(int32_t out4[4]);

/* ---- BVH: CudaAS / CudaBVH buffers (src/rt/cuda/CudaBVH.hpp:137-152) ----------------------- */
/* setBVH(CudaAS*): copy the three CudaBVH buffers (node, triWoop, triIndex) to the device.  layout = BVHLayout value
 * (CudaTracerKernels.hpp:52-63): Compact (4) and Compact2 (5) are traversed as they are; the basic layouts AOS_AOS (0),
 * AOS_SOA (1), SOA_AOS (2), SOA_SOA (3) of createNodeBasic / createTriWoopBasic / createTriIndexBasic (CudaBVH.cpp:453-575)
 * are kept as given (nt_bvh_download returns them unchanged) and rewritten on the device into the Compact form the
 * traversal kernel reads; the tree is validated on the way and a malformed one is refused. */
int nt_bvh_upload(int layout, const void* nodes, size_t nodeBytes,
                  const void* woop, size_t woopBytes,
                  const int32_t* triIndex, size_t idxBytes);
/* allocate an empty device BVH of the given sizes (replica side of a broadcast). */
int nt_bvh_alloc(int layout, size_t nodeBytes, size_t woopBytes, size_t idxBytes);
/* HLBVHBuilder(scene, platform, HLBVHParams) (src/rt/bvh/HLBVH/HLBVHBuilder.hpp:25-41):
 * GPU LBVH/HLBVH build straight into the Compact traversal layout.  outGpuSeconds = CUDA-event
 * time of the whole pipeline (reference: m_gpuTime, HLBVHBuilder.cpp:571). */
int nt_bvh_build(int builder, const float* vtxPos, int numVerts,
                 const int32_t* triVtxIndex, int numTris,
                 const float bboxLo[3], const float bboxHi[3],
                 int hlbvhBits, int leafSize, float epsilon,
                 float* outGpuSeconds);
/* Leaf formation of the GPU builder (NEW).  mode 0 (default) = the reference's rule: a child
 * range with <= leafSize triangles is a leaf (this is the mode every parity check runs in).  mode 1 = SAH-guided
 * collapse: the tree is emitted down to single triangles and subtrees are folded back into leaves of at most
 * maxLeafSize triangles (0 = leafSize) wherever that does not increase the SAH cost (Platform costs Cn = Ct = 1). */
int nt_bvh_set_collapse(int mode, int maxLeafSize);
/* Layout nt_bvh_build emits (NEW): BVHLayout_Compact (4, default; what the reference's HLBVHBuilder emits,
 * HLBVHBuilder.cpp:33) or BVHLayout_Compact2 (5).  Compact stores inner-child links as 32-bit BYTE offsets, which caps the
 * node buffer below the EntrypointSentinel 0x76543210 (1.98 GB = 31 M inner nodes); Compact2 stores offset / 16
 * (createCompact(bvh, 16), CudaBVH.cpp:86,614) and lifts that to 496 M nodes.  Trace a Compact2 BVH with
 * "kepler_dynamic_fetch" / "b200_persistent_speculative_while_while_compact2". */
int nt_bvh_set_build_layout(int layout);
/* Make the resident BVH BVHLayout_Compact (4) or BVHLayout_Compact2 (5) in place: AOS/SOA uploads are replaced by their
 * Compact form (CudaBVH.cpp:579-664 semantics: implicit leaves, terminator-delimited Woop lists), Compact <-> Compact2
 * rescales the inner-child offsets (createCompact's nodeOffsetSizeDiv, CudaBVH.cpp:86,614). */
int nt_bvh_convert(int layout);
/* NEW (no reference counterpart; the reference's flattening, CudaBVH::createCompact, CudaBVH.cpp:579-664, is host code too):
 * the 4-wide, 8-bit quantised node array ("Wide4", 64 bytes per node, format in csrc/nt_wide.cu) that the "b200_wide4*"
 * kernels traverse, derived from a Compact (4) / Compact2 (5) node buffer.  The library does this itself when such a kernel
 * is selected; this entry point exposes the same conversion on HOST buffers (no CUDA device needed) for tools and tests.
 * Leaves are shared with the source BVH: a Wide4 link < 0 is the source's ~woopIndex, so woop / triIndex are used as they are.
 * outWideNodes may be NULL to query *outWideBytes; *outMaxDepth (optional) = depth of the Wide4 tree. */
int nt_bvh_wide4_convert_host(int layout, const void* nodes, size_t nodeBytes, size_t woopBytes,
                              void* outWideNodes, size_t outCapacityBytes, size_t* outWideBytes, int* outMaxDepth);
/* NEW: the Wide4 node array of the RESIDENT BVH as the library derived it ON THE DEVICE (csrc/nt_wide.cu, wide4_convert_kernel: the
 * conversion the "b200_wide4*" / "b200_auto" kernels actually trace; derived here if no such kernel has run yet), copied to a host
 * or device buffer.  Same nodes, plane bytes and child-slot order as nt_bvh_wide4_convert_host produces from the same tree; only the
 * numbering of the nodes differs.  outWideNodes may be NULL to query *outWideBytes. */
int nt_bvh_wide4_download(void* outWideNodes, size_t outCapacityBytes, size_t* outWideBytes, int* outMaxDepth);
/* SAH cost of the resident BVH as the reference reports it in BVH::Stats (BVH.cpp:67-70: BVHNode::computeSubtreeProbabilities,
 * BVHNode.cpp:79-94, Platform costs 1 / 1: an inner node costs 2, a leaf its triangle count, each weighted by area / root area), computed
 * on the device in one parallel pass (the reference's GPU twin, calcSAH, is a single-thread recursion with inner cost 1,
 * emitTreeKernel.cu:1361-1400).  Also returns the inner-node, leaf and triangle-reference counts (any of the three may be NULL). */
int nt_bvh_sah(double* outSah, int64_t* outNumInner, int64_t* outNumLeaves, int64_t* outNumTris);
/* FW::hashBuffer (src/framework/base/Hash.cpp:33-75): the hash Renderer::getCudaBVH builds its cache file name from
 * ("bvhcache/<hash>_<builder>.dat", Renderer.cpp:173-178).  Pure host code (no device needed), so that both hosts above this ABI
 * name cache files with one implementation, pinned against the reference's own Hash.cpp (tests/test_reference_pin.py). */
int nt_hash_buffer(const void* ptr, size_t size, uint32_t* outHash);
/* NEW: generation of the resident BVH.  The library keeps ONE resident BVH (the reference's tracer keeps one CudaAS pointer,
 * CudaBVHTracer.hpp:46); every upload / alloc / build / convert replaces it and bumps this counter (0 = none).  A host-side handle
 * that skips the upload because "its" BVH is resident records the value after its build and compares it before tracing. */
int nt_bvh_generation(uint64_t* outGeneration);
/* sizes[3] = bytes of (nodes, woop, triIndex); layout of the resident BVH in *outLayout. */
int nt_bvh_sizes(size_t sizes[3], int* outLayout);
/* CudaBVH::serialize source buffers (CudaBVH.cpp:105-125): copy the device BVH out (host or device dst). */
int nt_bvh_download(void* nodes, void* woop, int32_t* triIndex);
/* device addresses of the three buffers, for the host's NCCL broadcast (NEW). */
int nt_bvh_device_ptrs(void* ptrs[3]);
/* ---- multi-GPU (NEW; SURVEY.md 8b/8e: the reference is single-context, CudaModule.hpp:92-97) ------------------------------
 * One process per GPU.  NCCL is bound at run time (dlopen of libnccl.so.2, or the path in NTRACE_NCCL_LIB): a single-GPU host needs
 * no NCCL.  Bootstrap: rank 0 calls nt_comm_unique_id and ships the 128 bytes to the other ranks by any means (file, socket, MPI);
 * every rank then calls nt_comm_init (collective; the communicator lives on the device nt_init selected).
 * nt_bvh_broadcast (collective) replicates the resident BVH of rank `root` — a header {layout, 3 sizes} and the three CudaBVH buffers,
 * ncclBroadcast over NVLink — into every other rank's library, which allocates as nt_bvh_alloc would; outSeconds = device time of the
 * three buffer broadcasts on this rank.  SURVEY.md sketched nt_bvh_broadcast(int numGpus); with one process per GPU the rank count
 * belongs to the communicator, and the ray path needs no numGpus argument: each rank passes ITS slice of a RayBuffer to nt_trace_batch.
 * nt_comm_allreduce: element-wise sum (op 0) / max (op 1) of `count` host doubles over the ranks (timing the slowest rank, counting rays). */
int nt_comm_unique_id(void* out128);
int nt_comm_init(int numRanks, int rank, const void* uniqueId128);
int nt_comm_destroy(void);
int nt_comm_allreduce(double* values, int count, int op);
int nt_bvh_broadcast(int root, float* outSeconds);
/* intermediate products of the last nt_bvh_build, for parity tests: sorted Morton keys and the
 * triangle order (reference buffers triMorton / triIdx, HLBVHBuilder.cpp:497-508). */
int nt_bvh_build_debug(uint32_t* sortedKeys, int32_t* sortedIdx, int numTris);

/* ---- trace: CudaBVHTracer::traceBatch(RayBuffer&) (CudaBVHTracer.cpp:88-168) --------------- */
/* rays: N x {float3 origin, float tmin, float3 dir, float tmax} (src/rt/Util.hpp:62-71);
 * results: N x {int id, float t, float u, float v} (Util.hpp:77-87; kernels store int4, CudaTracerKernels.hpp:222).
 * needClosestHit == 0 -> any-hit.  outSeconds = CUDA-event time around the kernel only. */
int nt_trace_batch(const float* rays, int32_t* results, int numRays, int needClosestHit,
                   float* outSeconds);
/* NEW (the reference traces one RayBuffer per launch, CudaBVHTracer.cpp:88-168, because its RayBuffer holds one batch; VERDICT round 1
 * item 4 asked for this form): numBatches device-resident batches traced by ONE persistent launch (groups of 64), every batch with its own
 * ray and result buffer and the same closest / any-hit flag.  Same kernels, same results per ray as numBatches calls of nt_trace_batch;
 * what goes away is the ramp-up and drain of every launch but one (one launch over the 24 x 1 Mi diffuse rays of the benchmark frame:
 * 3 673 Mrays/s; 24 launches: 2 851; 24 launches overlapped on two streams by nt_set_deferred(2): 3 489).  Device buffers only (rays
 * 32-byte, results 16-byte aligned); kernels: b200_persistent_speculative_while_while*, b200_wide4*, b200_auto*.  outSeconds = GPU time
 * of the launch(es); with nt_set_deferred(1 | 2) the call only enqueues (outSeconds = 0), ordered behind everything queued before it. */
int nt_trace_batches(int numBatches, const float* const* rays, int32_t* const* results, const int32_t* numRays, int needClosestHit,
                     float* outSeconds);
/* Asynchronous form of traceBatch (NEW; the reference's call is synchronous): submit a batch into one of 4 slots and
 * collect it later, so a host loop can keep independent batches of a frame in flight.  Buffers must be device memory or
 * pinned host memory and must stay valid until nt_trace_wait(slot) returns.  Host rays are DMA'd in on a copy stream,
 * results DMA'd out on another: with two or more slots in flight copy-in, traversal and copy-out of consecutive batches
 * overlap.  nt_trace_wait returns the kernel seconds of that batch like nt_trace_batch; waiting on an idle slot is a no-op. */
int nt_trace_batch_async(const float* rays, int32_t* results, int numRays, int needClosestHit, int slot);
int nt_trace_wait(int slot, float* outSeconds);

/* ---- ray generation: FW::RayGen (src/rt/ray/RayGen.cpp) ------------------------------------ */
/* RayGen::primary (RayGen.cpp:45-74): slot i -> pixel PixelTable[i]; idToSlot / slotToID may be NULL. */
int nt_raygen_primary(float* rays, int32_t* idToSlot, int32_t* slotToID,
                      const float origin[3], const float nscreenToWorld[16],
                      int w, int h, float maxDist, uint32_t randomSeed);
/* RayGen::ao (RayGen.cpp:198-232): numSamples rays per input slot in [first, first+numInputRays);
 * diffuse = same call with maxDist = camera far and closest-hit tracing (Renderer.cpp:533-538). */
int nt_raygen_ao(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID,
                 const float* inRays, const int32_t* inResults, const float* triNormals,
                 int firstInputSlot, int numInputRays, int numSamples,
                 float maxDist, uint32_t randomSeed);
/* NEW (the reference's only reordering is RayBuffer::mortonSort, a CPU sort of 192-bit keys, RayBuffer.cpp:103-163): slot order of the
 * rays nt_raygen_ao writes.  0 (default) = the reference kernel's order, slot == id (RayGenKernels.cu:129-236).  1 = the same rays, but
 * inside every tile of <= 1024 consecutive outputs (neighbouring hit points x their samples) slots are handed out by direction cell
 * (stable counting sort on a 16 x 16 octahedral map of the direction), idToSlot / slotToID carrying the permutation as mortonSort's do:
 * the warps of the trace kernel then fetch rays that leave neighbouring points in one direction.  Costs no extra pass over the rays. */
int nt_raygen_set_order(int mode);
/* RayGen::shadow + rayGenShadowKernel (RayGen.cpp:114-147, RayGenKernels.cu:240-302): numSamples any-hit rays from each
 * input hit point (backed off 1e-2 along the ray) towards a point in the cube of half-size lightRadius around lightPos;
 * tmax = distance to that point, or -1 (degenerate) when the input ray missed.  Used by the reference's VPL mode. */
int nt_raygen_shadow(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID,
                     const float* inRays, const int32_t* inResults,
                     int firstInputSlot, int numInputRays, int numSamples,
                     const float lightPos[3], float lightRadius, uint32_t randomSeed);
/* RayBuffer::mortonSort (src/rt/ray/RayBuffer.cpp:103-163): reorder the batch in place by the Morton key of
 * (origin, direction) and rebuild the id<->slot maps (outSlotToID[new] = inSlotToID[old], outIDToSlot[id] = new).
 * The reference sorts the 192-bit keys on the CPU (compareMortonKey, hash[5] most significant); here the same full-key order is
 * produced on the GPU by a word-wise stable LSD radix sort; rays with identical keys keep their original order. */
int nt_ray_sort(float* rays, int32_t* idToSlot, int32_t* slotToID, int numRays);
/* countHitsKernel (RendererKernels.cu:174-224, Renderer.cpp:693-705): results with id >= 0. */
int nt_count_hits(const int32_t* results, int numRays, int* outHits);
/* Scene::triNormal (src/rt/Scene.cpp:112): normalize(cross(v1 - v0, v2 - v0)) per triangle. */
int nt_tri_normals(const float* vtxPos, int numVerts, const int32_t* triVtxIndex, int numTris,
                   float* outNormals);

#ifdef __cplusplus
}
#endif
#endif /* NTRACE_B200_H */
