// ring_main.cpp -- Decoder -> VPP hand-over the way the reference's wrapper does it:
//   producer thread : publishes decoded NV12 surfaces (a device-resident synthetic pool here, NVDEC in the reference)
//                     into a FrameRing, paced at --fps or free-running
//   consumer threads: the unchanged getFrame sequence  ring.GetFrame(index, name, &frame); vpp.Convert(&frame, &out, params, name)
//                     (reference src/Wrappers/WrapperPython.cpp:296-312), one consumer name per thread
// Convert calls that are in flight at the same time are converted by ONE fused launch (tsvpp_convert_coalesced).
//
//   ring_main semantics                         GetFrame index rules of reference src/Decoder.cpp:97-131
//   ring_main run <consumers> <frames_per_consumer> <verify 0|1> [fps] [index]
//                                               prints aggregate frames/s, mean batch size, and (verify) checks every
//                                               output against a one-frame-per-launch conversion of the same surface.
//                                               fps 0 = the producer runs ahead of the consumers (2 M publishes/s).
//                                               Results are released the way the reference's callers do, cudaFree()
//                                               (verify) -- or, RING_FREE=async, stream-ordered with cudaFreeAsync():
//                                               cudaFree synchronises the device and serialises in the driver, which
//                                               alone caps ANY per-frame pipeline near 1e5 frames/s.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "FrameRing.h"
#include "VideoProcessor.h"
#include "tsvpp.h"

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            std::fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_));         \
            std::exit(3);                                                           \
        }                                                                           \
    } while (0)

static int fail(const char *what) {
    std::printf("FAIL %s\n", what);
    return 1;
}

static AVFrame surface(uint8_t *y, uint8_t *uv, int w, int h, int pitch) {
    AVFrame f;
    std::memset(&f, 0, sizeof(f));
    f.data[0] = y;
    f.data[1] = uv;
    f.linesize[0] = pitch;
    f.linesize[1] = pitch;
    f.width = w;
    f.height = h;
    return f;
}

static int semantics() {
    FrameRing ring(4);
    uint8_t *base = nullptr;
    CK(cudaMalloc(&base, 64 * 16));
    auto pub = [&](int k) { ring.Publish(surface(base + 64 * k, base + 64 * k + 32, 8, 2, 8)); };
    AVFrame out;
    // a consumer that arrives after frames exist gets the newest one at once
    pub(0);
    pub(1);
    if (ring.GetFrame(0, "a", &out) != 2 || out.data[0] != base + 64) return fail("index 0 = newest");
    // index > 0 is clamped to 0; the consumer must first see a NEW frame
    pub(2);
    if (ring.GetFrame(5, "a", &out) != 3 || out.data[0] != base + 128) return fail("index > 0 clamps to 0");
    pub(3);
    // negative index: older surfaces of the buffer, (currentFrame-1) % deep + index without wrap-around
    if (ring.GetFrame(-2, "a", &out) != 4 || out.data[0] != base + 64) return fail("index -2");
    pub(4);  // currentFrame = 5 -> slot (5-1) % 4 = 0
    if (ring.GetFrame(-1, "a", &out) != VREADER_REPEAT) return fail("slot 0 - 1 < 0 must be VREADER_REPEAT (no wrap)");
    pub(5);
    if (ring.GetFrame(-1, "a", &out) != 6 || out.data[0] != base + 64 * 4) return fail("index -1 after wrap of the ring");
    // a second consumer has its own flag; GetFrame blocks until something newer than its last frame arrives
    if (ring.GetFrame(0, "b", &out) != 6) return fail("second consumer sees the newest frame");
    std::atomic<int> got{0};
    std::thread waiter([&] {
        AVFrame o;
        got.store(ring.GetFrame(0, "b", &o));
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    if (got.load() != 0) return fail("GetFrame must block until a new frame is published");
    pub(6);
    waiter.join();
    if (got.load() != 7) return fail("blocked consumer woken by Publish");
    // end of stream: waiting consumers are released with the reference's exception
    std::atomic<int> threw{0};
    std::thread waiter2([&] {
        AVFrame o;
        try {
            ring.GetFrame(0, "b", &o);
        } catch (const std::runtime_error &) {
            threw.store(1);
        }
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    ring.notifyConsumers();
    waiter2.join();
    if (!threw.load()) return fail("finished stream must throw 'Decoding finished'");
    cudaFree(base);
    std::printf("semantics ok\n");
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && std::string(argv[1]) == "semantics") return semantics();
    if (argc < 5 || std::string(argv[1]) != "run") {
        std::fprintf(stderr, "usage: ring_main semantics | run <consumers> <frames_per_consumer> <verify> [fps] [index]\n");
        return 2;
    }
    const int consumers = std::atoi(argv[2]), per = std::atoi(argv[3]), verify = std::atoi(argv[4]);
    const double fps = argc > 5 ? std::atof(argv[5]) : 0.0;
    const int index = argc > 6 ? std::atoi(argv[6]) : 0;
    const bool free_async = std::getenv("RING_FREE") && std::string(std::getenv("RING_FREE")) == "async";
    // RING_OUT=pool: every consumer recycles 32 result buffers of its own through VideoProcessor::ConvertInto (what the
    // ATen hand-off does with torch's caching allocator) -- no allocation, no free, no CUDA call per frame
    const bool pool_out = std::getenv("RING_OUT") && std::string(std::getenv("RING_OUT")) == "pool";
    // RING_SOURCE=direct: every consumer walks the surface pool by itself (one decoder per consumer, never paced by
    // another thread) -- isolates what the Convert side can take from what one shared ring can hand out
    const bool direct_src = std::getenv("RING_SOURCE") && std::string(std::getenv("RING_SOURCE")) == "direct";
    const int kPoolOut = 32;
    const int W = 1920, H = 1080, pitch = 1920, DW = 256, DH = 256, pool = 64;
    // C2: BGR24 planar 256x256 bilinear, normalised float32
    ColorOptions color(BGR24);
    color.planesPos = PLANAR;
    color.normalization = true;
    ResizeOptions resize(DW, DH);
    resize.type = BILINEAR;
    const size_t out_elems = (size_t)3 * DW * DH;

    // synthetic decoded surfaces, device resident
    std::vector<uint8_t *> ys(pool), uvs(pool);
    std::vector<uint8_t> host((size_t)pitch * H * 3 / 2);
    uint32_t lcg = 12345u;
    for (int k = 0; k < pool; ++k) {
        for (auto &b : host) {
            lcg = lcg * 1664525u + 1013904223u;
            b = (uint8_t)(lcg >> 24);
        }
        CK(cudaMalloc(&ys[k], (size_t)pitch * H));
        CK(cudaMalloc(&uvs[k], (size_t)pitch * H / 2));
        CK(cudaMemcpy(ys[k], host.data(), (size_t)pitch * H, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(uvs[k], host.data() + (size_t)pitch * H, (size_t)pitch * H / 2, cudaMemcpyHostToDevice));
    }

    auto logger = std::make_shared<Logger>();
    VideoProcessor vpp;
    if (vpp.Init(logger, (uint8_t)consumers, false) != VREADER_OK) return fail("Init");

    // expected result of every surface: one frame per launch through the plain C ABI
    std::vector<std::vector<float>> expect;
    if (verify) {
        TsvppContext *ctx = nullptr;
        if (tsvpp_create(0, &ctx)) return fail("tsvpp_create");
        TsvppParams p;
        std::memset(&p, 0, sizeof(p));
        p.dst_width = DW; p.dst_height = DH; p.resize_type = TSVPP_BILINEAR; p.fourcc = TSVPP_BGR24;
        p.planes = TSVPP_PLANAR; p.normalization = 1;
        float *d = nullptr;
        CK(cudaMalloc(&d, out_elems * sizeof(float)));
        expect.resize(pool, std::vector<float>(out_elems));
        for (int k = 0; k < pool; ++k) {
            TsvppFrame f = {ys[k], uvs[k], pitch, pitch, W, H};
            void *outs[1] = {d};
            if (tsvpp_convert_batch(ctx, &f, 1, &p, outs, nullptr)) return fail("reference conversion");
            CK(cudaMemcpy(expect[k].data(), d, out_elems * sizeof(float), cudaMemcpyDeviceToHost));
        }
        cudaFree(d);
        tsvpp_destroy(ctx);
    }

    FrameRing ring(10);
    std::atomic<bool> stop{false};
    std::atomic<unsigned long long> published{0};
    std::thread producer([&] {
        cudaSetDevice(0);
        using clk = std::chrono::steady_clock;
        const auto t0 = clk::now();
        for (unsigned long long n = 0; !stop.load(std::memory_order_relaxed); ++n) {
            const int k = (int)(n % pool);
            ring.Publish(surface(ys[k], uvs[k], W, H, pitch));
            published.store(n + 1, std::memory_order_relaxed);
            // paced like a decoder; "free-running" still leaves the ring's mutex to the consumers most of the time
            const double rate = fps > 0 ? fps : 2e6;
            const auto due = t0 + std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>((n + 1) / rate));
            while (clk::now() < due && !stop.load(std::memory_order_relaxed)) {
                if (fps > 0 && fps < 1e5) std::this_thread::yield();
            }
        }
    });

    std::atomic<int> bad{0}, repeats{0};
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int c = 0; c < consumers; ++c)
        th.emplace_back([&, c] {
            cudaSetDevice(0);
            const std::string name = "consumer" + std::to_string(c);
            FrameParameters params = {resize, color, CropOptions()};
            std::vector<void *> held;
            std::vector<int> held_src;
            std::vector<void *> mine;
            if (pool_out) {
                mine.resize(kPoolOut);
                for (auto &b : mine) CK(cudaMalloc(&b, out_elems * sizeof(float)));
            }
            std::vector<float> back(verify ? out_elems : 0);
            auto flush = [&] {
                if (pool_out) vpp.WaitCoalesced();  // before the buffers are read back / written again
                for (size_t i = 0; i < held.size(); ++i) {
                    if (verify) {
                        CK(cudaMemcpy(back.data(), held[i], out_elems * sizeof(float), cudaMemcpyDeviceToHost));
                        if (std::memcmp(back.data(), expect[held_src[i]].data(), out_elems * sizeof(float)) != 0) bad.fetch_add(1);
                    }
                    if (pool_out)
                        continue;  // recycled
                    if (free_async)
                        CK(cudaFreeAsync(held[i], 0));  // stream-ordered: the legacy stream is ordered after the blocking coalescing stream
                    else
                        CK(cudaFree(held[i]));  // the caller owns the result and releases it with cudaFree, as with the reference
                }
                held.clear();
                held_src.clear();
            };
            for (int n = 0; n < per; ++n) {
                AVFrame in, out;
                std::memset(&out, 0, sizeof(out));
                int src = -1;
                if (direct_src) {
                    src = (c * 7 + n) % pool;
                    in = surface(ys[src], uvs[src], W, H, pitch);
                } else {
                    const int sts = ring.GetFrame(index, name, &in);
                    if (sts == VREADER_REPEAT) {
                        repeats.fetch_add(1);
                        --n;
                        continue;
                    }
                    for (int k = 0; k < pool && src < 0; ++k)
                        if (in.data[0] == ys[k]) src = k;
                }
                const int rc = pool_out ? vpp.ConvertInto(&in, &out, mine[held.size()], params, name)
                                        : vpp.Convert(&in, &out, params, name);
                if (rc != VREADER_OK || !out.opaque || out.width != DW || out.height != DH) {
                    bad.fetch_add(1000);
                    break;
                }
                held.push_back(out.opaque);
                held_src.push_back(src);
                if (held.size() >= (pool_out ? (size_t)kPoolOut : free_async ? 16u : 64u)) flush();
            }
            flush();
            for (auto &b : mine) cudaFree(b);
        });
    for (auto &t : th) t.join();
    CK(cudaDeviceSynchronize());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    stop.store(true);
    ring.notifyConsumers();
    producer.join();
    const double total = (double)consumers * per;
    std::printf("consumers %d frames %.0f seconds %.4f fps %.0f published %llu repeats %d bad %d coalesce %s async %s free %s\n", consumers,
                total, secs, total / secs, published.load(), repeats.load(), bad.load(),
                std::getenv("TSVPP_COALESCE") ? std::getenv("TSVPP_COALESCE") : "1",
                std::getenv("TSVPP_ASYNC") ? std::getenv("TSVPP_ASYNC") : "0", pool_out ? "recycled(ConvertInto)" : free_async ? "cudaFreeAsync" : "cudaFree");
    if (direct_src) std::printf("source direct (no shared ring)\n");
    vpp.Close();
    return bad.load() ? 1 : 0;
}
