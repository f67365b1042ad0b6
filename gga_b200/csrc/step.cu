// One training-shaped step of the geometry hot path from HOST buffers, entirely behind the C
// ABI: H2D copies, membership (index build + streaming), projection + loss forward/backward,
// D2H copies of masks, loss and gradients — what a reference-side caller holding numpy /
// CPU-tensor data would invoke (the `points_in_boxes_cpu`-style contract of
// /root/reference/mmdet3d/ops/__init__.py:12,38 extended to the whole step of
// mmdet3d/models/dense_heads/centerpoint_head_gga.py:629-723).
//
// The PCIe link is the bound here (46 MB per step at the training shape), so the step is
// pipelined frame by frame over a few streams: the H2D copy of frame f+1, the kernels of
// frame f and the D2H copy of the masks of frame f-1 overlap (full-duplex link); the box
// kernel and its small copies run on their own stream.  A context owns all device buffers,
// streams and events, so a call performs no allocation.
#include <new>

#include "common.cuh"

namespace {

constexpr int kMaxStreams = 8;

struct StepCtx {
  int F, N, M, pts_stride, W, device, n_streams;
  float* d_points;    // [F, N, stride]
  float* d_boxes;     // [F, M, 7]
  float* d_proj;      // [F, M, 16]
  float* d_target;    // [F, M, 4]
  float* d_weight;    // [F, M]
  uint32_t* d_bits;   // [F, N, W]
  float* d_box2d;     // [F*M, 4]
  float* d_loss;      // [F*M]
  float* d_loss_sum;  // [1]
  float* d_grad;      // [F*M, 7]
  void* d_scratch;    // loss-reduction scratch of this context (zeroed at creation)
  size_t scratch_bytes;
  int32_t* d_pairs;   // [pair_capacity, 2] hit list of the masks (allocated on first use)
  int32_t* d_count;   // [1]
  int32_t* h_count;   // [1] page-locked: the number of pairs of the step in flight
  int pair_capacity;
  bool pending;       // a submitted step has not been waited for
  int pending_hits;   // its hit capacity (0: dense rows or masks left on the device)
  cudaStream_t streams[kMaxStreams];
  cudaStream_t box_stream;
  cudaEvent_t boxes_ready, done[kMaxStreams], box_done;
};

void destroy(StepCtx* c) {
  if (!c) return;
  if (c->pending) {  // host buffers of a step in flight must not be referenced after destroy
    cudaStreamSynchronize(c->box_stream);
    for (int i = 0; i < c->n_streams; ++i) cudaStreamSynchronize(c->streams[i]);
  }
  cudaFree(c->d_points); cudaFree(c->d_boxes); cudaFree(c->d_proj); cudaFree(c->d_target); cudaFree(c->d_weight);
  cudaFree(c->d_bits); cudaFree(c->d_box2d); cudaFree(c->d_loss); cudaFree(c->d_loss_sum); cudaFree(c->d_grad);
  cudaFree(c->d_scratch);
  if (c->d_pairs) cudaFree(c->d_pairs);
  if (c->d_count) cudaFree(c->d_count);
  if (c->h_count) cudaFreeHost(c->h_count);
  for (int i = 0; i < kMaxStreams; ++i) {
    if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
  }
  if (c->box_stream) cudaStreamDestroy(c->box_stream);
  if (c->boxes_ready) cudaEventDestroy(c->boxes_ready);
  if (c->box_done) cudaEventDestroy(c->box_done);
  delete c;
}

}  // namespace

extern "C" int gga_step_create(int num_frames, int num_points, int num_boxes, int pts_stride, int n_streams,
                               void** ctx_out) {
  GGA_REQUIRE(ctx_out != nullptr, "null ctx_out");
  *ctx_out = nullptr;
  GGA_REQUIRE(num_frames > 0 && num_points > 0 && num_boxes > 0, "sizes must be positive");
  GGA_REQUIRE(pts_stride >= 3, "pts_stride must be >= 3 (got %d)", pts_stride);
  if (n_streams < 1) n_streams = 3;
  if (n_streams > kMaxStreams) n_streams = kMaxStreams;
  StepCtx* c = new (std::nothrow) StepCtx();
  GGA_REQUIRE(c != nullptr, "out of host memory");
  c->F = num_frames; c->N = num_points; c->M = num_boxes; c->pts_stride = pts_stride;
  c->W = gga_pib_row_words(num_boxes);
  c->n_streams = n_streams;
  const size_t F = num_frames, N = num_points, M = num_boxes;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_points, F * N * pts_stride * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_boxes, F * M * 7 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_proj, F * M * 16 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_target, F * M * 4 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_weight, F * M * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_bits, F * N * c->W * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_box2d, F * M * 4 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_loss, F * M * 4 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_loss_sum, sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_grad, F * M * 7 * sizeof(float));
  c->scratch_bytes = gga_loss_scratch_bytes();
  if (e == cudaSuccess) e = cudaMalloc(&c->d_scratch, c->scratch_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->d_scratch, 0, c->scratch_bytes);
  for (int i = 0; i < n_streams && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->box_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->boxes_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->box_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    gga_set_error("gga_step_create: %s", cudaGetErrorString(e));
    destroy(c);
    return GGA_ERR_CUDA;
  }
  *ctx_out = c;
  return GGA_OK;
}

extern "C" int gga_step_destroy(void* ctx) {
  destroy(static_cast<StepCtx*>(ctx));
  return GGA_OK;
}

// Enqueues the whole step; returns on the first failure WITHOUT synchronising (the caller does).
static int enqueue_step(StepCtx* c, const float* points, const float* boxes, const float* lidar2img,
                        const float* target, const float* weight, int proj_mode, int loss_kind, float loss_weight,
                        float avg_factor, float eps, float depth_clamp, uint32_t* bits, bool hits, float* loss_sum,
                        float* grad_boxes) {
  const size_t F = c->F, N = c->N, M = c->M, st = c->pts_stride, W = c->W;
  if (hits) GGA_CHECK_CUDA(cudaMemsetAsync(c->d_count, 0, sizeof(int32_t), c->box_stream));
  // boxes first: every frame's membership needs them
  GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_boxes, boxes, F * M * 7 * sizeof(float), cudaMemcpyHostToDevice, c->box_stream));
  GGA_CHECK_CUDA(cudaEventRecord(c->boxes_ready, c->box_stream));
  if (c->n_streams == 1) {
    // One piece: a single copy each way and ONE membership launch over all frames.  Nothing overlaps
    // inside the step — meant for callers that keep several steps in flight (gga_step_submit_host on
    // alternating contexts), where the copies of different steps overlap and fewer, larger copies win.
    cudaStream_t q = c->streams[0];
    GGA_CHECK_CUDA(cudaStreamWaitEvent(q, c->boxes_ready, 0));
    GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_points, points, F * N * st * sizeof(float), cudaMemcpyHostToDevice, q));
    const int rc = gga_points_in_boxes_bits(c->d_points, (int)st, c->d_boxes, c->d_bits, (int)F, (int)N, (int)M, q);
    if (rc != GGA_OK) return rc;
    if (bits)
      GGA_CHECK_CUDA(cudaMemcpyAsync(bits, c->d_bits, F * N * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, q));
    if (hits) {
      const int rc2 = gga_pib_hit_list(c->d_bits, (int64_t)(F * N), (int)M, 0, c->d_pairs, c->pair_capacity, c->d_count, 0, q);
      if (rc2 != GGA_OK) return rc2;
    }
    GGA_CHECK_CUDA(cudaEventRecord(c->done[0], q));
  }
  // membership, one frame per pipeline slot
  for (size_t f = 0; f < F && c->n_streams > 1; ++f) {
    const int s = (int)(f % c->n_streams);
    cudaStream_t q = c->streams[s];
    GGA_CHECK_CUDA(cudaStreamWaitEvent(q, c->boxes_ready, 0));
    GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_points + f * N * st, points + f * N * st, N * st * sizeof(float),
                                   cudaMemcpyHostToDevice, q));
    const int rc = gga_points_in_boxes_bits(c->d_points + f * N * st, (int)st, c->d_boxes + f * M * 7,
                                            c->d_bits + f * N * W, 1, (int)N, (int)M, q);
    if (rc != GGA_OK) return rc;
    if (bits)
      GGA_CHECK_CUDA(cudaMemcpyAsync(bits + f * N * W, c->d_bits + f * N * W, N * W * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, q));
    if (hits) {
      const int rc2 = gga_pib_hit_list(c->d_bits + f * N * W, (int64_t)N, (int)M, (int64_t)(f * N), c->d_pairs,
                                       c->pair_capacity, c->d_count, 0, q);
      if (rc2 != GGA_OK) return rc2;
    }
    GGA_CHECK_CUDA(cudaEventRecord(c->done[s], q));
  }
  // projection + loss forward / backward
  cudaStream_t b = c->box_stream;
  GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_proj, lidar2img, F * M * 16 * sizeof(float), cudaMemcpyHostToDevice, b));
  GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_target, target, F * M * 4 * sizeof(float), cudaMemcpyHostToDevice, b));
  if (weight) GGA_CHECK_CUDA(cudaMemcpyAsync(c->d_weight, weight, F * M * sizeof(float), cudaMemcpyHostToDevice, b));
  gga_box_loss_args a = {};
  a.boxes = c->d_boxes; a.proj = c->d_proj; a.proj_stride = 16;
  a.target = c->d_target;
  a.weight = weight ? c->d_weight : nullptr; a.weight_cols = 1;
  a.n = (int)(F * M); a.mode = proj_mode; a.loss_kind = loss_kind;
  a.depth_clamp = depth_clamp; a.eps = eps; a.grad_scale = loss_weight / avg_factor;
  a.box2d = c->d_box2d; a.loss = c->d_loss; a.loss_sum = c->d_loss_sum; a.grad_boxes = c->d_grad;
  a.scratch = c->d_scratch; a.scratch_bytes = c->scratch_bytes;
  const int rc = gga_box_project_loss(&a, b);
  if (rc != GGA_OK) return rc;
  GGA_CHECK_CUDA(cudaMemcpyAsync(grad_boxes, c->d_grad, F * M * 7 * sizeof(float), cudaMemcpyDeviceToHost, b));
  GGA_CHECK_CUDA(cudaMemcpyAsync(loss_sum, c->d_loss_sum, sizeof(float), cudaMemcpyDeviceToHost, b));
  return GGA_OK;
}

// Drains every stream of the context; returns the first CUDA error seen.
static cudaError_t drain(StepCtx* c) {
  cudaError_t e = cudaStreamSynchronize(c->box_stream);
  for (int s = 0; s < c->n_streams; ++s) {
    const cudaError_t es = cudaStreamSynchronize(c->streams[s]);
    if (e == cudaSuccess) e = es;
  }
  return e;
}

extern "C" int gga_step_submit_host(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                                    const float* target, const float* weight, int proj_mode, int loss_kind,
                                    float loss_weight, float avg_factor, float eps, float depth_clamp,
                                    uint32_t* bits, int hit_capacity, float* loss_sum, float* grad_boxes) {
  StepCtx* c = static_cast<StepCtx*>(ctx);
  GGA_REQUIRE(c != nullptr, "null context");
  GGA_REQUIRE(points && boxes && lidar2img && target && loss_sum && grad_boxes, "null host pointer");
  GGA_REQUIRE(avg_factor > 0.f && hit_capacity >= 0, "avg_factor must be positive, hit_capacity non-negative");
  GGA_REQUIRE(!c->pending, "the context already has a step in flight: call gga_step_wait_host first");
  int dev = 0;
  GGA_CHECK_CUDA(cudaGetDevice(&dev));
  GGA_REQUIRE(dev == c->device, "context belongs to device %d, current device is %d", c->device, dev);
  const bool hits = hit_capacity > 0;
  if (hits && c->pair_capacity < hit_capacity) {  // (re)allocated outside the pipeline, first use only
    GGA_CHECK_CUDA(cudaDeviceSynchronize());
    if (c->d_pairs) cudaFree(c->d_pairs);
    c->d_pairs = nullptr;
    c->pair_capacity = 0;
    GGA_CHECK_CUDA(cudaMalloc(&c->d_pairs, (size_t)hit_capacity * 2 * sizeof(int32_t)));
    if (!c->d_count) GGA_CHECK_CUDA(cudaMalloc(&c->d_count, sizeof(int32_t)));
    if (!c->h_count) GGA_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->h_count), sizeof(int32_t), cudaHostAllocDefault));
    c->pair_capacity = hit_capacity;
  }
  int rc = enqueue_step(c, points, boxes, lidar2img, target, weight, proj_mode, loss_kind, loss_weight, avg_factor,
                        eps, depth_clamp, bits, hits, loss_sum, grad_boxes);
  if (rc == GGA_OK && hits) {  // the number of pairs, once every frame's list kernel is done
    cudaStream_t b = c->box_stream;
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < c->n_streams && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(b, c->done[s], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_count, c->d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, b);
    if (e != cudaSuccess) {
      gga_set_error("gga_step_submit_host: %s", cudaGetErrorString(e));
      rc = GGA_ERR_CUDA;
    }
  }
  if (rc != GGA_OK) {
    // on a failure half-way the copies already enqueued still reference the caller's host buffers
    drain(c);
    return rc;
  }
  c->pending = true;
  c->pending_hits = hits ? hit_capacity : 0;
  return GGA_OK;
}

extern "C" int gga_step_wait_host(void* ctx, int32_t* hits, int32_t* n_hits) {
  StepCtx* c = static_cast<StepCtx*>(ctx);
  GGA_REQUIRE(c != nullptr, "null context");
  GGA_REQUIRE(c->pending, "no step in flight");
  c->pending = false;
  cudaError_t e = drain(c);
  if (e != cudaSuccess) {
    gga_set_error("gga_step_wait_host: %s", cudaGetErrorString(e));
    return GGA_ERR_CUDA;
  }
  if (c->pending_hits > 0) {
    GGA_REQUIRE(hits != nullptr && n_hits != nullptr, "the step was submitted with a hit list: hits and n_hits are required");
    const int cap = c->pending_hits;
    *n_hits = *c->h_count;
    const int n = *n_hits < cap ? *n_hits : cap;
    if (n > 0) {  // exactly that many pairs
      e = cudaMemcpyAsync(hits, c->d_pairs, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->box_stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->box_stream);
      if (e != cudaSuccess) {
        gga_set_error("gga_step_wait_host: %s", cudaGetErrorString(e));
        return GGA_ERR_CUDA;
      }
    }
    if (*n_hits > cap) {
      gga_set_error("hit list overflow: %d pairs, capacity %d", *n_hits, cap);
      return GGA_ERR_UNSUPPORTED;
    }
  }
  return GGA_OK;
}

// The synchronous forms, like the CPU op they stand in for: submit + wait.
extern "C" int gga_step_run_host(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                                 const float* target, const float* weight, int proj_mode, int loss_kind,
                                 float loss_weight, float avg_factor, float eps, float depth_clamp,
                                 uint32_t* bits, float* loss_sum, float* grad_boxes) {
  const int rc = gga_step_submit_host(ctx, points, boxes, lidar2img, target, weight, proj_mode, loss_kind, loss_weight,
                                      avg_factor, eps, depth_clamp, bits, 0, loss_sum, grad_boxes);
  return rc != GGA_OK ? rc : gga_step_wait_host(ctx, nullptr, nullptr);
}

extern "C" int gga_step_run_host_hits(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                                      const float* target, const float* weight, int proj_mode, int loss_kind,
                                      float loss_weight, float avg_factor, float eps, float depth_clamp,
                                      int32_t* hits, int hit_capacity, int32_t* n_hits, float* loss_sum,
                                      float* grad_boxes) {
  GGA_REQUIRE(hits && n_hits, "null host pointer");
  GGA_REQUIRE(hit_capacity > 0, "hit_capacity must be positive");
  const int rc = gga_step_submit_host(ctx, points, boxes, lidar2img, target, weight, proj_mode, loss_kind, loss_weight,
                                      avg_factor, eps, depth_clamp, nullptr, hit_capacity, loss_sum, grad_boxes);
  return rc != GGA_OK ? rc : gga_step_wait_host(ctx, hits, n_hits);
}

extern "C" int gga_step_device_bits(void* ctx, uint32_t** bits_device) {
  StepCtx* c = static_cast<StepCtx*>(ctx);
  GGA_REQUIRE(c != nullptr && bits_device != nullptr, "null argument");
  *bits_device = c->d_bits;
  return GGA_OK;
}
