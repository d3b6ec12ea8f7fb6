// Feasibility probe: how fast does nvJPEG decode the frames of an MJPG AVI on this GPU, per backend?
// Parses the RIFF 'movi' list for '00dc' chunks, decodes up to 16 384 frames in batches of 512 to
// interleaved BGR (the reference decodes to BGR, features.py:226,235) and prints frames/s; also the
// mean absolute difference of the first frame's bytes against a raw BGR dump if one is given.
// Build: nvcc -O2 -std=c++17 -o scripts/microbench/bin/nvjpeg_probe scripts/microbench/nvjpeg_probe.cu -lnvjpeg
// Usage: nvjpeg_probe clip.avi [first_frame_bgr.raw]
#include <nvjpeg.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static std::vector<uint8_t> read_file(const char* path) {
    std::vector<uint8_t> buf;
    FILE* f = fopen(path, "rb");
    if (!f) return buf;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n);
    if (fread(buf.data(), 1, n, f) != static_cast<size_t>(n)) buf.clear();
    fclose(f);
    return buf;
}
static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (uint32_t(p[3]) << 24); }

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    const std::vector<uint8_t> file = read_file(argv[1]);
    if (file.size() < 12 || memcmp(file.data(), "RIFF", 4) || memcmp(file.data() + 8, "AVI ", 4)) {
        printf("not an AVI file\n");
        return 1;
    }
    // walk the top-level chunks; descend into LIST 'movi'
    std::vector<const uint8_t*> ptrs;
    std::vector<size_t> lens;
    size_t pos = 12;
    while (pos + 8 <= file.size()) {
        const uint32_t size = rd32(file.data() + pos + 4);
        if (!memcmp(file.data() + pos, "LIST", 4) && !memcmp(file.data() + pos + 8, "movi", 4)) {
            size_t q = pos + 12;
            const size_t end = pos + 8 + size;
            while (q + 8 <= end && q + 8 <= file.size()) {
                const uint32_t csz = rd32(file.data() + q + 4);
                if (!memcmp(file.data() + q + 2, "dc", 2) && csz > 4) {
                    ptrs.push_back(file.data() + q + 8);
                    lens.push_back(csz);
                }
                q += 8 + csz + (csz & 1);
            }
        }
        pos += 8 + size + (size & 1);
    }
    printf("frames in movi: %zu, first frame %zu bytes, markers %02x%02x\n", ptrs.size(),
           lens.empty() ? 0 : lens[0], ptrs.empty() ? 0 : ptrs[0][0], ptrs.empty() ? 0 : ptrs[0][1]);
    if (ptrs.empty()) return 1;
    const int total = ptrs.size() < 16384 ? static_cast<int>(ptrs.size()) : 16384;
    const int batch = 512;
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    const char* names[] = {"DEFAULT", "HYBRID", "GPU_HYBRID", "HARDWARE"};
    const nvjpegBackend_t backends[] = {NVJPEG_BACKEND_DEFAULT, NVJPEG_BACKEND_HYBRID,
                                        NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_HARDWARE};
    for (int b = 0; b < 4; ++b) {
        nvjpegHandle_t h;
        nvjpegStatus_t st = nvjpegCreateEx(backends[b], nullptr, nullptr, 0, &h);
        if (st != NVJPEG_STATUS_SUCCESS) {
            printf("%-10s create failed: status %d\n", names[b], st);
            continue;
        }
        nvjpegJpegState_t state;
        nvjpegJpegStateCreate(h, &state);
        int comps = 0, widths[4] = {0}, heights[4] = {0};
        nvjpegChromaSubsampling_t sub;
        st = nvjpegGetImageInfo(h, ptrs[0], lens[0], &comps, &sub, widths, heights);
        if (st != NVJPEG_STATUS_SUCCESS) {
            printf("%-10s GetImageInfo failed: status %d\n", names[b], st);
            continue;
        }
        const int W = widths[0], H = heights[0];
        if (b == 0) printf("image %d x %d, %d components, subsampling enum %d\n", W, H, comps, sub);
        st = nvjpegDecodeBatchedInitialize(h, state, batch, 1, NVJPEG_OUTPUT_BGRI);
        if (st != NVJPEG_STATUS_SUCCESS) {
            printf("%-10s DecodeBatchedInitialize failed: status %d\n", names[b], st);
            continue;
        }
        uint8_t* out_dev = nullptr;
        cudaMalloc(&out_dev, static_cast<size_t>(batch) * W * H * 3);
        std::vector<nvjpegImage_t> outs(batch);
        for (int i = 0; i < batch; ++i) {
            memset(&outs[i], 0, sizeof(nvjpegImage_t));
            outs[i].channel[0] = out_dev + static_cast<size_t>(i) * W * H * 3;
            outs[i].pitch[0] = static_cast<size_t>(W) * 3;
        }
        bool ok = true;
        double secs = 0;
        for (int rep = 0; rep < 2 && ok; ++rep) {   // rep 0 warms up
            const auto t0 = std::chrono::steady_clock::now();
            for (int i0 = 0; i0 + batch <= total && ok; i0 += batch) {
                st = nvjpegDecodeBatched(h, state, ptrs.data() + i0, lens.data() + i0, outs.data(), stream);
                if (st != NVJPEG_STATUS_SUCCESS) {
                    printf("%-10s DecodeBatched failed at frame %d: status %d\n", names[b], i0, st);
                    ok = false;
                }
            }
            cudaStreamSynchronize(stream);
            secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (ok) {
            const int done = total / batch * batch;
            printf("%-10s %d frames in %.3f s = %.0f frames/s\n", names[b], done, secs, done / secs);
            if (argc > 2) {   // first frame of the LAST batch vs the raw dump of frame `done - batch`
                const std::vector<uint8_t> want = read_file(argv[2]);
                std::vector<uint8_t> got(static_cast<size_t>(W) * H * 3);
                // decode frame 0 alone for the comparison
                nvjpegDecodeBatched(h, state, ptrs.data(), lens.data(), outs.data(), stream);
                cudaStreamSynchronize(stream);
                cudaMemcpy(got.data(), out_dev, got.size(), cudaMemcpyDeviceToHost);
                if (want.size() == got.size()) {
                    double sum = 0;
                    int mx = 0;
                    for (size_t i = 0; i < got.size(); ++i) {
                        const int d = abs(int(got[i]) - int(want[i]));
                        sum += d;
                        mx = d > mx ? d : mx;
                    }
                    printf("%-10s frame 0 vs cv2: mean |diff| %.4f, max %d\n", names[b], sum / got.size(), mx);
                }
            }
        }
        cudaFree(out_dev);
        nvjpegJpegStateDestroy(state);
        nvjpegDestroy(h);
    }
    return 0;
}
