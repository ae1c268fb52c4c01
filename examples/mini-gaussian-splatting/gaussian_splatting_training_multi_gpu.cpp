// examples/mini-gaussian-splatting/gaussian_splatting_training_multi_gpu.cpp -- the training loop of the reference's
// GaussianSplattingTrainer::train() (examples/mini-gaussian-splatting/gaussian_splatting_training.cu:110-183) on the
// GPUs of one NVSwitch box, host C++ on libxyz_b200.so only (no Python, no framework): ONE host thread per GPU.
//
// The reference is single-GPU, one target image.  Here the Gaussians are replicated and the work of an iteration is sharded
//   --mode views   V target views round-robin over the G GPUs (BASELINE configs[4]: 3 M Gaussians, 8 views, 8 GPUs);
//                  view v = the reference's test image (image_utils.cpp:59-75) rolled by (64 v, 37 v) pixels
//   --mode rows    ONE image, tile-aligned row bands (BASELINE configs[3] on G GPUs)
// and per iteration every GPU runs   loss = 0 ; launch (forward + backward of its share) ; exchange + Adam   where
//   --exchange peer          xyz_adam_step_individual_peer: reduce-scatter + Adam + all-gather + zero-grad + loss
//                            all-reduce as ONE kernel over NVLink peer memory (default; with --graph the whole iteration
//                            is ONE captured CUDA graph replayed max-iterations times: no host work per iteration)
//   --exchange nccl          xyz_allreduce_grads (NCCL all-reduce of N x 9 floats + the loss) + Adam on every replica
//   --exchange nccl-sharded  xyz_adam_step_individual_sharded (NCCL reduce-scatter, Adam on the range, all-gather)
// Every launch uses a caller-owned workspace (xyz_launch_gaussian_splatting_ws): nothing allocates or synchronises
// inside the loop; the host reads the loss every --report iterations only.
//
//   gaussian_splatting_training_multi_gpu [--gpus G] [--views V] [--mode views|rows] [--exchange peer|nccl|nccl-sharded]
//        [--graph] [--num-gaussians N] [--image WxH] [--max-iterations I] [--report K] [--seed S]
//        [--lr c s r col o] [--precise]
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <xyz_b200.h>

namespace {

struct Options {
    int gpus = 1, views = 1, n = 1000, w = 256, h = 256, iterations = 100, report = 10;
    unsigned seed = 42;
    std::string mode = "views", exchange = "peer";
    bool graph = false;
    int flags = 0;
    float lr[5] = {0.5f, 0.01f, 0.01f, 0.01f, 0.02f};
};

#define CUDA_OK(expr)                                                                                              \
    do {                                                                                                           \
        const cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define XYZ_OK(expr)                                                                                               \
    do {                                                                                                           \
        const int rc_ = (expr);                                                                                    \
        if (rc_ != 0) throw std::runtime_error(std::string(#expr) + " failed with " + std::to_string(rc_));        \
    } while (0)

class Barrier {  // C++17 has no std::barrier
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen != gen_; });
        }
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

std::vector<float> test_image(int w, int h, int view) {  // create_test_image, rolled per view
    std::vector<float> img(static_cast<size_t>(w) * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int sx = ((x - 64 * view) % w + w) % w, sy = ((y - 37 * view) % h + h) % h;
            float* p = &img[(static_cast<size_t>(y) * w + x) * 3];
            p[0] = static_cast<float>(sx) / w;
            p[1] = static_cast<float>(sy) / h;
            p[2] = 0.5f * (p[0] + p[1]);
        }
    return img;
}

std::vector<xyz_gaussian_params> initialize_random(int n, int w, int h, unsigned seed) {  // gaussian_parameters.cu:27-66
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> pos_x(0.0f, static_cast<float>(w)), pos_y(0.0f, static_cast<float>(h));
    std::uniform_real_distribution<float> color(0.1f, 0.2f), opacity(0.05f, 0.1f), scale(0.0f, 2.0f);
    std::vector<xyz_gaussian_params> g(static_cast<size_t>(n));
    for (auto& p : g) {
        p.center[0] = pos_x(rng);
        p.center[1] = pos_y(rng);
        p.scale[0] = std::max(1.0f, scale(rng));
        p.scale[1] = std::max(1.0f, scale(rng));
        p.rotation[0] = 0.0f;
        p.color[0] = color(rng);
        p.color[1] = color(rng);
        p.color[2] = color(rng);
        p.opacity[0] = opacity(rng);
    }
    return g;
}

struct Shared {  // what the rank threads exchange through host memory (one process: device pointers are valid everywhere)
    Options opt;
    std::vector<void*> mailbox, params, grads;
    std::vector<xyz_comm*> comms;
    std::vector<std::string> errors;
    std::vector<double> ms_per_iter;
    std::vector<float> losses;  // rank 0: total loss at every report
    Barrier* barrier = nullptr;
};

void rank_main(int rank, Shared* sh) {
    const Options& o = sh->opt;
    const int G = o.gpus, N = o.n, W = o.w, H = o.h;
    try {
        CUDA_OK(cudaSetDevice(rank));
        for (int p = 0; p < G; ++p)
            if (p != rank) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(e);
                cudaGetLastError();
            }
        cudaStream_t st;
        CUDA_OK(cudaStreamCreate(&st));
        // this rank's share
        std::vector<int> my_views;
        int row_begin = 0, row_end = H;
        if (o.mode == "rows") {
            const int tiles_y = (H + 15) / 16;
            row_begin = std::min(H, tiles_y * rank / G * 16);
            row_end = std::min(H, tiles_y * (rank + 1) / G * 16);
            my_views.push_back(0);
        } else {
            for (int v = rank; v < o.views; v += G) my_views.push_back(v);
        }
        // buffers: params / grads / mailbox in memory every rank can reach
        void *params = nullptr, *grads = nullptr, *mailbox = nullptr;
        XYZ_OK(xyz_peer_alloc(sizeof(xyz_gaussian_params) * N, &params, nullptr));
        XYZ_OK(xyz_peer_alloc(sizeof(xyz_gaussian_grads) * N, &grads, nullptr));
        unsigned char unused_handle[64];
        XYZ_OK(xyz_peer_mailbox_create(&mailbox, unused_handle));
        xyz_adam_state* adam = nullptr;
        CUDA_OK(cudaMalloc(&adam, sizeof(xyz_adam_state) * N));
        CUDA_OK(cudaMemset(adam, 0, sizeof(xyz_adam_state) * N));
        float* loss = nullptr;
        CUDA_OK(cudaMalloc(&loss, sizeof(float)));
        const std::vector<xyz_gaussian_params> host_params = initialize_random(N, W, H, o.seed);
        CUDA_OK(cudaMemcpy(params, host_params.data(), sizeof(xyz_gaussian_params) * N, cudaMemcpyHostToDevice));
        const size_t P = static_cast<size_t>(W) * H;
        std::vector<float*> targets, outputs;
        for (int v : my_views) {
            const std::vector<float> img = test_image(W, H, v);
            float *t = nullptr, *out = nullptr;
            CUDA_OK(cudaMalloc(&t, sizeof(float) * 3 * P));
            CUDA_OK(cudaMalloc(&out, sizeof(float) * 3 * P));
            CUDA_OK(cudaMemcpy(t, img.data(), sizeof(float) * 3 * P, cudaMemcpyHostToDevice));
            CUDA_OK(cudaMemset(out, 0, sizeof(float) * 3 * P));
            targets.push_back(t);
            outputs.push_back(out);
        }
        // workspace: learn this scene's list length from ONE ordinary launch, then size for 1.5 x of it
        void* ws = nullptr;
        size_t ws_bytes = 0;
        long long max_entries = 0;
        const bool has_work = !my_views.empty() && row_end > row_begin;
        if (has_work) {
            void* scratch_grads = nullptr;
            CUDA_OK(cudaMalloc(&scratch_grads, sizeof(xyz_gaussian_grads) * N));
            CUDA_OK(cudaMemset(scratch_grads, 0, sizeof(xyz_gaussian_grads) * N));
            CUDA_OK(cudaMemset(loss, 0, sizeof(float)));
            XYZ_OK(xyz_launch_gaussian_splatting_rows(static_cast<xyz_gaussian_params*>(params),
                                                      static_cast<xyz_gaussian_grads*>(scratch_grads), targets[0], outputs[0],
                                                      loss, W, H, N, row_begin, row_end, st, o.flags));
            long long stats[4];
            XYZ_OK(xyz_splat_last_stats(stats));
            max_entries = stats[0] + stats[0] / 2 + 4096;
            CUDA_OK(cudaFree(scratch_grads));
            XYZ_OK(xyz_b200_shutdown());  // the library-owned scratch of that one launch is not needed again
            ws_bytes = xyz_splat_workspace_bytes(W, H, N, row_begin, row_end, max_entries, o.flags);
            if (ws_bytes == 0) throw std::runtime_error("xyz_splat_workspace_bytes: unsupported shape");
            CUDA_OK(cudaMalloc(&ws, ws_bytes));
            XYZ_OK(xyz_splat_workspace_init(ws, ws_bytes, st));
        }
        sh->mailbox[rank] = mailbox;
        sh->params[rank] = params;
        sh->grads[rank] = grads;
        CUDA_OK(cudaDeviceSynchronize());
        sh->barrier->wait();
        xyz_peer_group group{};
        xyz_peer_splat_buffers bufs{};
        for (int p = 0; p < G; ++p) {
            group.mailbox[p] = sh->mailbox[p];
            bufs.params[p] = static_cast<xyz_gaussian_params*>(sh->params[p]);
            bufs.grads[p] = static_cast<xyz_gaussian_grads*>(sh->grads[p]);
        }
        group.rank = rank;
        group.world = G;
        xyz_comm* comm = (o.exchange != "peer" && G > 1) ? sh->comms[rank] : nullptr;

        auto enqueue_iteration = [&](int iteration /* 1-based; 0 = counted on the device */) {
            CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
            if (has_work)
                for (size_t k = 0; k < targets.size(); ++k)
                    XYZ_OK(xyz_launch_gaussian_splatting_ws(static_cast<xyz_gaussian_params*>(params),
                                                            static_cast<xyz_gaussian_grads*>(grads), targets[k], outputs[k], loss,
                                                            W, H, N, row_begin, row_end, ws, ws_bytes, max_entries, st, o.flags));
            if (o.exchange == "peer") {
                XYZ_OK(xyz_adam_step_individual_peer(&group, &bufs, adam, N, o.lr, 0.9f, 0.999f, 1e-8f, iteration, loss, st));
            } else if (o.exchange == "nccl-sharded" && G > 1) {
                XYZ_OK(xyz_adam_step_individual_sharded(comm, static_cast<xyz_gaussian_params*>(params),
                                                        static_cast<xyz_gaussian_grads*>(grads), adam, N, o.lr, 0.9f, 0.999f,
                                                        1e-8f, iteration, loss, st));
            } else {
                if (G > 1) {
                    XYZ_OK(xyz_allreduce_grads(comm, static_cast<float*>(grads), 9LL * N, st));
                    XYZ_OK(xyz_allreduce_grads(comm, loss, 1, st));
                }
                XYZ_OK(xyz_adam_step_individual_zero_grads(static_cast<xyz_gaussian_params*>(params),
                                                           static_cast<xyz_gaussian_grads*>(grads), adam, N, o.lr, 0.9f, 0.999f,
                                                           1e-8f, iteration, st));
            }
        };

        cudaGraphExec_t exec = nullptr;
        if (o.graph) {
            if (o.exchange != "peer") throw std::runtime_error("--graph needs --exchange peer (device-side Adam step counter)");
            cudaGraph_t graph = nullptr;
            CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_iteration(0);
            CUDA_OK(cudaStreamEndCapture(st, &graph));
            CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
            CUDA_OK(cudaGraphDestroy(graph));
        }
        CUDA_OK(cudaDeviceSynchronize());
        sh->barrier->wait();
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 1; it <= o.iterations; ++it) {
            if (exec) CUDA_OK(cudaGraphLaunch(exec, st));
            else enqueue_iteration(it);
            if (it == 1 || it % o.report == 0 || it == o.iterations) {
                float l = 0.f;  // the reference reads the loss EVERY iteration (:150-151); here only when it is printed
                CUDA_OK(cudaMemcpyAsync(&l, loss, sizeof(float), cudaMemcpyDeviceToHost, st));
                CUDA_OK(cudaStreamSynchronize(st));
                if (rank == 0) {
                    sh->losses.push_back(l);
                    const int images = o.mode == "rows" ? 1 : o.views;
                    std::printf("Iteration %4d | average Loss: %.6e\n", it, l / (static_cast<float>(W) * H * images));
                }
            }
        }
        CUDA_OK(cudaStreamSynchronize(st));
        sh->ms_per_iter[rank] =
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / o.iterations;
        if (has_work) {
            long long status[4];
            XYZ_OK(xyz_splat_workspace_status(ws, st, status));
            if (status[1] != 0) throw std::runtime_error("a launch did not fit its workspace (raise the head-room)");
        }
        sh->barrier->wait();  // nobody frees memory a peer may still touch
        if (exec) cudaGraphExecDestroy(exec);
        for (float* p : targets) cudaFree(p);
        for (float* p : outputs) cudaFree(p);
        cudaFree(ws);
        cudaFree(loss);
        cudaFree(adam);
        xyz_peer_mailbox_destroy(mailbox);
        xyz_peer_mailbox_destroy(params);
        xyz_peer_mailbox_destroy(grads);
        cudaStreamDestroy(st);
    } catch (const std::exception& e) {
        sh->errors[rank] = e.what();
        std::fprintf(stderr, "rank %d: %s\n", rank, e.what());
        std::exit(1);  // the other ranks would wait for this one forever
    }
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value after %s\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--gpus") o.gpus = std::atoi(next());
        else if (a == "--views") o.views = std::atoi(next());
        else if (a == "--mode") o.mode = next();
        else if (a == "--exchange") o.exchange = next();
        else if (a == "--graph") o.graph = true;
        else if (a == "--num-gaussians") o.n = std::atoi(next());
        else if (a == "--image") {
            if (std::sscanf(next(), "%dx%d", &o.w, &o.h) != 2) {
                std::fprintf(stderr, "expected --image WxH\n");
                return 2;
            }
        } else if (a == "--max-iterations") o.iterations = std::atoi(next());
        else if (a == "--report") o.report = std::atoi(next());
        else if (a == "--seed") o.seed = static_cast<unsigned>(std::atoll(next()));
        else if (a == "--precise") o.flags |= XYZ_FLAG_PRECISE_MATH;
        else if (a == "--lr") {
            for (int k = 0; k < 5; ++k) o.lr[k] = static_cast<float>(std::atof(next()));
        } else {
            std::fprintf(stderr, "Unknown argument: %s\n", a.c_str());
            return 2;
        }
    }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) {
        std::fprintf(stderr, "Error: no CUDA device\n");
        return 1;
    }
    if (o.gpus < 1 || o.gpus > have || o.gpus > XYZ_PEER_MAX_WORLD || o.views < 1 || o.n < 1 || o.w < 1 || o.h < 1 ||
        o.iterations < 1 || o.report < 1 || (o.mode != "views" && o.mode != "rows") ||
        (o.exchange != "peer" && o.exchange != "nccl" && o.exchange != "nccl-sharded")) {
        std::fprintf(stderr, "Error: invalid options (%d GPU(s) present)\n", have);
        return 2;
    }
    std::printf("%s\n=== Multi-GPU Gaussian Splatting Training ===\nGPUs: %d  mode: %s  views: %d  exchange: %s%s\n"
                "Image: %dx%d  Gaussians: %d  iterations: %d\n",
                xyz_b200_version(), o.gpus, o.mode.c_str(), o.mode == "rows" ? 1 : o.views, o.exchange.c_str(),
                o.graph ? " (one CUDA graph per iteration)" : "", o.w, o.h, o.n, o.iterations);
    Shared sh;
    sh.opt = o;
    sh.mailbox.assign(o.gpus, nullptr);
    sh.params.assign(o.gpus, nullptr);
    sh.grads.assign(o.gpus, nullptr);
    sh.comms.assign(o.gpus, nullptr);
    sh.errors.assign(o.gpus, "");
    sh.ms_per_iter.assign(o.gpus, 0.0);
    Barrier barrier(o.gpus);
    sh.barrier = &barrier;
    if (o.exchange != "peer" && o.gpus > 1) {
        const int rc = xyz_comm_init_all(sh.comms.data(), o.gpus, nullptr);
        if (rc != 0) {
            std::fprintf(stderr, "Error: xyz_comm_init_all failed with %d (is libnccl.so.2 on the library path?)\n", rc);
            return 1;
        }
    }
    std::vector<std::thread> threads;
    for (int r = 0; r < o.gpus; ++r) threads.emplace_back(rank_main, r, &sh);
    for (auto& t : threads) t.join();
    for (xyz_comm* c : sh.comms) xyz_comm_destroy(c);
    const double ms = *std::max_element(sh.ms_per_iter.begin(), sh.ms_per_iter.end());
    const float scale = static_cast<float>(o.w) * o.h * (o.mode == "rows" ? 1 : o.views);
    std::printf("Training completed!  first average loss %.6e, last %.6e, %.3f ms/iteration (max over ranks, host clock)\n",
                sh.losses.front() / scale, sh.losses.back() / scale, ms);
    return 0;
}
