// examples/mini-gaussian-splatting/gaussian_splatting_training.cpp -- host driver of the 2-D splat fit, on
// libxyz_b200.so.
//
// Role of the reference's GaussianSplattingTrainer (examples/mini-gaussian-splatting/
// gaussian_splatting_training.cu:15-233) + TrainingConfigParser (training_config.cpp:28-146) +
// GaussianCollection (gaussian_parameters.cu): random Gaussians (initialize_random, :27-66), then per iteration
//   zero_gradients_gpu -> total_loss = 0 -> launch_gaussian_splatting -> read total_loss -> adam_step_gpu_individual
// (here the zero-grad pass is fused into the Adam pass: xyz_adam_step_individual_zero_grads)
// (training loop :127-175).  The launch goes through include/xyz_b200_compat.hpp, i.e. the reference's own
// launch_gaussian_splatting(...) signature; zero-grad and Adam are the C-ABI replacements of the reference kernels.
//
// Command line = the reference's: five positional learning rates (center scale rotation color opacity), then
//   --target PATH --max-iterations N --save-interval N --num-gaussians N --no-save-images
//   --beta1 x --beta2 x --epsilon x --help                       (same validation rules, training_config.cpp:113-146)
// Differences: the image codec (stb, out of scope) is replaced by binary PPM (P6) for --target and for the saved
// renderings; `--target synthetic[:WxH]` uses the reference's create_test_image gradient pattern
// (image_utils.cpp:59-75).  Extra: --seed S (the reference seeds from the clock), --deterministic, --precise.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include <xyz_autodiff/util/cuda_unique_ptr.cuh>

// the reference's struct layouts (gaussian_parameters.h:12-41); PixelOutput = ConstArray<float, 3>
struct GaussianParams { float center[2], scale[2], rotation[1], color[3], opacity[1]; };
struct GaussianGrads { float center[2], scale[2], rotation[1], color[3], opacity[1]; };
namespace xyz_autodiff {
template <typename T, int N>
struct ConstArray { T data[N]; };
}  // namespace xyz_autodiff
static int g_launch_flags = 0;
#define XYZ_B200_COMPAT_FLAGS g_launch_flags
#include <xyz_b200_compat.hpp>
static_assert(sizeof(GaussianParams) == sizeof(xyz_gaussian_params), "layout");
static_assert(sizeof(PixelOutput) == 12, "layout");

namespace {

struct TrainingConfig {
    float lr_center = 0.01f, lr_scale = 0.01f, lr_rotation = 0.01f, lr_color = 0.01f, lr_opacity = 0.01f;
    int max_iterations = 500, save_interval = 25, num_gaussians = 1000;
    std::string target_image_path;
    bool save_images = true;
    float beta1 = 0.9f, beta2 = 0.999f, epsilon = 1e-8f;
    unsigned seed = 42;
};

void print_usage(const char* prog) {
    std::cout << "Usage: " << prog << " <lr_center> <lr_scale> <lr_rotation> <lr_color> <lr_opacity> [OPTIONS]\n"
              << "  --target PATH.ppm | synthetic[:WxH]   --max-iterations N   --save-interval N   --num-gaussians N\n"
              << "  --no-save-images   --beta1 x   --beta2 x   --epsilon x   --seed S   --deterministic   --precise\n";
}

TrainingConfig parse(int argc, char** argv) {
    TrainingConfig c;
    if (argc < 6) {
        print_usage(argv[0]);
        throw std::runtime_error("Insufficient arguments");
    }
    int i = 1;
    c.lr_center = static_cast<float>(std::atof(argv[i++]));
    c.lr_scale = static_cast<float>(std::atof(argv[i++]));
    c.lr_rotation = static_cast<float>(std::atof(argv[i++]));
    c.lr_color = static_cast<float>(std::atof(argv[i++]));
    c.lr_opacity = static_cast<float>(std::atof(argv[i++]));
    for (; i < argc; ++i) {
        const std::string a = argv[i];
        const bool has_value = i + 1 < argc;
        if (a == "--target" && has_value) c.target_image_path = argv[++i];
        else if (a == "--max-iterations" && has_value) c.max_iterations = std::atoi(argv[++i]);
        else if (a == "--save-interval" && has_value) c.save_interval = std::atoi(argv[++i]);
        else if (a == "--num-gaussians" && has_value) c.num_gaussians = std::atoi(argv[++i]);
        else if (a == "--no-save-images") c.save_images = false;
        else if (a == "--beta1" && has_value) c.beta1 = static_cast<float>(std::atof(argv[++i]));
        else if (a == "--beta2" && has_value) c.beta2 = static_cast<float>(std::atof(argv[++i]));
        else if (a == "--epsilon" && has_value) c.epsilon = static_cast<float>(std::atof(argv[++i]));
        else if (a == "--seed" && has_value) c.seed = static_cast<unsigned>(std::atoll(argv[++i]));
        else if (a == "--deterministic") g_launch_flags |= XYZ_FLAG_DETERMINISTIC;
        else if (a == "--precise") g_launch_flags |= XYZ_FLAG_PRECISE_MATH;
        else if (a == "--help") {
            print_usage(argv[0]);
            std::exit(0);
        } else {
            std::cerr << "Unknown argument: " << a << std::endl;
            print_usage(argv[0]);
            throw std::runtime_error("Invalid argument");
        }
    }
    if (c.lr_center <= 0 || c.lr_scale <= 0 || c.lr_rotation <= 0 || c.lr_color <= 0 || c.lr_opacity <= 0)
        throw std::runtime_error("All learning rates must be positive");
    if (c.max_iterations <= 0) throw std::runtime_error("Max iterations must be positive");
    if (c.save_interval <= 0) throw std::runtime_error("Save interval must be positive");
    if (c.num_gaussians <= 0) throw std::runtime_error("Number of Gaussians must be positive");
    if (c.target_image_path.empty()) throw std::runtime_error("Target image path must be specified with --target");
    if (c.beta1 <= 0 || c.beta1 >= 1) throw std::runtime_error("Beta1 must be in range (0, 1)");
    if (c.beta2 <= 0 || c.beta2 >= 1) throw std::runtime_error("Beta2 must be in range (0, 1)");
    if (c.epsilon <= 0) throw std::runtime_error("Epsilon must be positive");
    return c;
}

struct ImageData {
    int width = 0, height = 0;
    std::vector<float> rgb;  // H x W x 3 in [0, 1]
};

ImageData synthetic_image(int w, int h) {  // create_test_image: r = x/W, g = y/H, b = (r + g) / 2
    ImageData img{w, h, std::vector<float>(static_cast<size_t>(w) * h * 3)};
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float* p = &img.rgb[(static_cast<size_t>(y) * w + x) * 3];
            p[0] = static_cast<float>(x) / w;
            p[1] = static_cast<float>(y) / h;
            p[2] = 0.5f * (p[0] + p[1]);
        }
    return img;
}

ImageData load_ppm(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Failed to load image: " + path);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    f >> magic >> w >> h >> maxv;
    f.get();
    if (magic != "P6" || w <= 0 || h <= 0 || maxv != 255) throw std::runtime_error("Unsupported PPM (need binary P6, 8 bit): " + path);
    std::vector<unsigned char> raw(static_cast<size_t>(w) * h * 3);
    f.read(reinterpret_cast<char*>(raw.data()), static_cast<std::streamsize>(raw.size()));
    if (!f) throw std::runtime_error("Truncated PPM: " + path);
    ImageData img{w, h, std::vector<float>(raw.size())};
    for (size_t i = 0; i < raw.size(); ++i) img.rgb[i] = raw[i] / 255.0f;  // u8 -> float / 255 like load_image
    return img;
}

void save_ppm(const std::string& path, const ImageData& img) {
    std::ofstream f(path, std::ios::binary);
    f << "P6\n" << img.width << " " << img.height << "\n255\n";
    std::vector<unsigned char> raw(img.rgb.size());
    for (size_t i = 0; i < raw.size(); ++i)
        raw[i] = static_cast<unsigned char>(std::min(255.0f, std::max(0.0f, img.rgb[i] * 255.0f)));
    f.write(reinterpret_cast<const char*>(raw.data()), static_cast<std::streamsize>(raw.size()));
}

ImageData load_target(const std::string& spec) {
    if (spec.rfind("synthetic", 0) == 0) {
        int w = 256, h = 256;
        const size_t colon = spec.find(':');
        if (colon != std::string::npos && std::sscanf(spec.c_str() + colon + 1, "%dx%d", &w, &h) != 2)
            throw std::runtime_error("expected synthetic:WxH");
        return synthetic_image(w, h);
    }
    return load_ppm(spec);
}

std::vector<GaussianParams> initialize_random(int n, int w, int h, std::mt19937& rng) {
    std::uniform_real_distribution<float> pos_x(0.0f, static_cast<float>(w)), pos_y(0.0f, static_cast<float>(h));
    std::uniform_real_distribution<float> color(0.1f, 0.2f), opacity(0.05f, 0.1f), scale(0.0f, 2.0f);
    std::vector<GaussianParams> g(static_cast<size_t>(n));
    for (auto& p : g) {  // same draw order as the reference: x, y, s0, s1, r, g, b, opacity
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

#define XYZ_CALL(expr)                                                                           \
    do {                                                                                         \
        const int rc_ = (expr);                                                                  \
        if (rc_ != 0) throw std::runtime_error(std::string(#expr) + " failed with " + std::to_string(rc_)); \
    } while (0)

}  // namespace

int main(int argc, char** argv) {
    try {
        const TrainingConfig config = parse(argc, argv);
        CHECK_CUDA_ERROR(cudaSetDevice(0));
        const ImageData target = load_target(config.target_image_path);
        const int W = target.width, H = target.height, N = config.num_gaussians;
        std::cout << xyz_b200_version() << "\n=== Starting Gaussian Splatting Training ===\n"
                  << "Target image: " << W << "x" << H << "\nGaussians: " << N << "\nLearning rates: center "
                  << config.lr_center << ", scale " << config.lr_scale << ", rotation " << config.lr_rotation
                  << ", color " << config.lr_color << ", opacity " << config.lr_opacity << "\nMax iterations: "
                  << config.max_iterations << std::endl;

        std::mt19937 rng(config.seed);
        const std::vector<GaussianParams> host_params = initialize_random(N, W, H, rng);
        const size_t P = static_cast<size_t>(W) * H;
        auto d_params = makeCudaUniqueArray<GaussianParams>(N);
        auto d_grads = makeCudaUniqueArray<GaussianGrads>(N);
        auto d_adam = makeCudaUniqueArray<xyz_adam_state>(N);
        auto d_target = makeCudaUniqueArray<PixelOutput>(P);
        auto d_output = makeCudaUniqueArray<PixelOutput>(P);
        auto d_loss = makeCudaUnique<float>();
        CHECK_CUDA_ERROR(cudaMemcpy(d_params.get(), host_params.data(), sizeof(GaussianParams) * N, cudaMemcpyHostToDevice));
        CHECK_CUDA_ERROR(cudaMemset(d_adam.get(), 0, sizeof(xyz_adam_state) * N));
        CHECK_CUDA_ERROR(cudaMemcpy(d_target.get(), target.rgb.data(), sizeof(float) * 3 * P, cudaMemcpyHostToDevice));
        CHECK_CUDA_ERROR(cudaMemset(d_output.get(), 0, sizeof(float) * 3 * P));

        auto save_rendering = [&](int iteration) {
            ImageData img{W, H, std::vector<float>(P * 3)};
            CHECK_CUDA_ERROR(cudaMemcpy(img.rgb.data(), d_output.get(), sizeof(float) * 3 * P, cudaMemcpyDeviceToHost));
            char name[64];
            std::snprintf(name, sizeof(name), "output/iteration_%04d.ppm", iteration);
            save_ppm(name, img);
        };
        if (config.save_images) {
            if (std::system("mkdir -p output") != 0) std::cerr << "[warning]: failed to make directory";
            save_ppm("output/target.ppm", target);
        }

        const float lr[5] = {config.lr_center, config.lr_scale, config.lr_rotation, config.lr_color, config.lr_opacity};
        float first_loss = 0.f, last_loss = 0.f;
        double total_ms = 0.0;
        // zero_gradients_gpu(): once up front; afterwards the Adam pass clears every gradient it has consumed
        XYZ_CALL(xyz_zero_gradients(reinterpret_cast<xyz_gaussian_grads*>(d_grads.get()), N, nullptr));
        for (int iteration = 0; iteration < config.max_iterations; ++iteration) {
            const auto t0 = std::chrono::high_resolution_clock::now();
            CHECK_CUDA_ERROR(cudaMemsetAsync(d_loss.get(), 0, sizeof(float), nullptr));
            launch_gaussian_splatting(d_params.get(), d_grads.get(), d_target.get(), d_output.get(), d_loss.get(), W, H, N);
            float total_loss = 0.0f;  // the reference's only synchronisation point (:150-151)
            CHECK_CUDA_ERROR(cudaMemcpy(&total_loss, d_loss.get(), sizeof(float), cudaMemcpyDeviceToHost));
            XYZ_CALL(xyz_adam_step_individual_zero_grads(reinterpret_cast<xyz_gaussian_params*>(d_params.get()),
                                                         reinterpret_cast<xyz_gaussian_grads*>(d_grads.get()), d_adam.get(), N,
                                                         lr, config.beta1, config.beta2, config.epsilon, iteration + 1, nullptr));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
            total_ms += ms;
            const float average_loss = total_loss / (static_cast<float>(H) * W);
            if (iteration == 0) first_loss = average_loss;
            last_loss = average_loss;
            if (iteration % 10 == 0)
                std::cout << "Iteration " << std::setw(4) << iteration << " | average Loss: " << std::scientific
                          << std::setprecision(6) << average_loss << " | Time: " << std::fixed << std::setprecision(3) << ms
                          << "ms" << std::endl;
            if (config.save_images && iteration % config.save_interval == 0) save_rendering(iteration);
        }
        CHECK_CUDA_ERROR(cudaDeviceSynchronize());
        if (config.save_images) save_rendering(config.max_iterations);
        std::cout << "Training completed!  first average loss " << std::scientific << first_loss << ", last " << last_loss
                  << ", " << std::fixed << std::setprecision(3) << total_ms / config.max_iterations << " ms/iteration"
                  << std::endl;
        xyz_b200_shutdown();
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
