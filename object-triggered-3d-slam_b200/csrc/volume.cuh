// volume.cuh -- host-side state of one ScalableTSDFVolume replacement (see volume.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace otslam {

constexpr int kMaxBatch = 32;   // frames fused per block residency (one bit each in the entry mask)

// per-frame constants consumed by the allocation (FP64 pose) and integration (FP32 extrinsic) kernels
struct FrameDev {
    float E[12];      // rows 0..2 of extrinsic.cast<float>()           
    float es[3];      // (E * voxel_length).col(2): the per-z increment  
    int32_t key_off;  // multi-object arenas: x-key offset of the frame's object (0 for ordinary volumes)
    double pose[12];  // rows 0..2 of camera_pose = extrinsic.inverse()  
};
static_assert(sizeof(FrameDev) == 160, "FrameDev layout");

constexpr int kNB = 3;          // batch buffers: allocation runs up to two batches ahead of integration
// device counters: [0] pool slots handed out; per batch buffer b: [2+b] work-list length, [5+b] flags
enum CounterIdx { kPoolCount = 0, kListCount = 2, kFlags = 5, kNumCounters = 8 };
enum Flags { kFlagHashFull = 1, kFlagKeyRange = 2, kFlagBoxTooLarge = 4 };

struct MeshResult {
    int64_t nv = 0, nf = 0;
    double* d_verts = nullptr;    // [nv][3]
    double* d_colors = nullptr;   // [nv][3]
    double* d_normals = nullptr;  // [nv][3]
    int32_t* d_faces = nullptr;   // [nf][3]
    int32_t* d_ekeys = nullptr;   // [nv][4]
    void release();
};
struct PointsResult {
    int64_t n = 0;
    double* d_pts = nullptr;
    double* d_cols = nullptr;
    int32_t* d_ekeys = nullptr;
    void release();
};

}  // namespace otslam

struct otslam_volume {
    int device = 0;
    double voxel_length = 0, sdf_trunc = 0, unit_length = 0;
    int color_type = OTSLAM_COLOR_RGB8;
    otslam::SlabSpec slab{0, 8, 1, 0, 1};
    cudaStream_t stream = nullptr, copy_stream = nullptr, pre_stream = nullptr;
    bool own_stream = true;
    int batch = otslam::kMaxBatch;
    int zsplit = 0;                // CTAs per block along z in the integration kernel: 0 = default (2), 1, 2, 4, 8
    int n_objects = 0;             // > 0: multi-object arena (object id in the block key), see common.cuh
    int sel_obj = -1;              // arena: object that extraction / export / stats address (-1 = none selected)
    int64_t obj_frames[otslam::kMaxObjects] = {};   // frames integrated per object (the 65535 limit is per voxel)
    int64_t frames_integrated = 0;

    // block hash: open addressing, entry index is the handle used by the per-batch work list
    uint32_t cap = 0;
    uint64_t* d_keys = nullptr;
    int32_t* d_vals = nullptr;     // entry -> pool slot
    // kNB-buffered per batch: allocation of batches b+1, b+2 runs on pre_stream while batch b integrates
    uint32_t* d_masks[otslam::kNB] = {};   // entry -> bit f set when frame f of the batch touches it
    int32_t* d_list[otslam::kNB] = {};     // entries touched by the batch (each once)
    int32_t* d_order[otslam::kNB] = {};    // the same entries, most-frames-first (launch order of the integration CTAs)
    uint32_t* d_lmask[otslam::kNB] = {};   // their frame masks in that order (the hash-side masks are cleared when this is built)

    // block pool: chunks of kChunkBlocks blocks, 64 KiB per block
    std::vector<uint4*> chunks;
    uint4** d_chunks = nullptr;
    int64_t n_blocks = 0;          // host mirror of the pool counter

    int* d_counters = nullptr;
    int* h_counters = nullptr;     // pinned [kNB][kNumCounters]
    std::vector<uint4*> h_chunk_table;   // never reallocates (reserved to kMaxChunks): async copies read from it
    cudaEvent_t ev_pre_done[otslam::kNB] = {}, ev_k4_done[otslam::kNB] = {}, ev_main = nullptr;

    // frame staging (kNB buffers)
    uint16_t* d_raw_depth[otslam::kNB] = {};
    uint8_t* d_raw_rgb[otslam::kNB] = {};
    uint2* d_packed[otslam::kNB] = {};
    size_t raw_frames_cap = 0, raw_px_cap = 0, packed_cap = 0, packed_px = 0;
    size_t raw_depth_bytes_per_px = 2;
    cudaEvent_t ev_copied[otslam::kNB] = {}, ev_raw_free[otslam::kNB] = {};
    otslam::FrameDev* d_frames[otslam::kNB] = {};
    otslam::FrameDev* h_frames = nullptr;  // pinned [kNB][kMaxBatch]

    // depth->camera-distance multiplier image, cached per intrinsics
    float* d_mult = nullptr;
    int mult_w = 0, mult_h = 0;
    double mult_intr[4] = {0, 0, 0, 0};

    // optional per-kernel timing (otslam_volume_profile)
    bool profiling = false;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_launches[4] = {0, 0, 0, 0};
    std::vector<cudaEvent_t> prof_events;          // pairs, recycled
    std::vector<int> prof_pending;                 // kernel id per recorded pair
    size_t prof_used = 0;

    otslam::MeshResult mesh;
    otslam::PointsResult points;

    // device-resident list of the allocated blocks sorted by key (extraction order, export, halo selection); rebuilt
    // when `epoch` (bumped by integrate / reset / halo import) has moved past `sorted_epoch`
    uint64_t epoch = 1, sorted_epoch = 0;
    uint64_t* d_sorted_keys = nullptr;
    int32_t* d_sorted_slots = nullptr;
    int64_t n_sorted = 0;

    // halo pieces packed by otslam_volume_halo_pack: grouped by destination rank, (key, kind) order inside a group
    int32_t* d_halo_keys = nullptr;      // [n_halo][4]
    uint4* d_halo_planes = nullptr;      // [n_halo][256]
    int64_t n_halo = 0;
    std::vector<int64_t> halo_counts;    // pieces per destination rank

    cudaEvent_t ev_ext = nullptr;        // otslam_volume_wait_stream
};

namespace otslam {
// implemented in volume.cu, used by extract.cu: the allocated blocks sorted by key, in HBM (valid until the volume changes)
int volume_sorted_blocks_device(otslam_volume* v, const uint64_t** d_keys, const int32_t** d_slots, int* n);
// the same restricted to the selected object of a multi-object arena (x_off = the object's x-key offset; 0 otherwise)
int volume_selected_blocks_device(otslam_volume* v, const uint64_t** d_keys, const int32_t** d_slots, int* n, int* x_off);
// the same downloaded to the host (export_blocks)
int volume_sorted_blocks(otslam_volume* v, std::vector<uint64_t>& keys, std::vector<int32_t>& slots);
// knn.cu: stable LSD radix sort of (u64 key, i32 value) pairs by the low `bits` bits; result in (kb, vb), (ka, va) clobbered
int device_sort_pairs(uint64_t* ka, int32_t* va, uint64_t* kb, int32_t* vb, int64_t n, int bits, cudaStream_t s);
}
