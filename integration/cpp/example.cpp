// C++ user of include/yume_pdq_cuda.hpp, modelled on the reference's integration/dummy.c and
// examples/end_to_end.rs: hash a frame (8 dihedrals), plant one of them in a database, find it.
// Build: g++ -std=c++17 -Iinclude integration/cpp/example.cpp -Lyume_pdq_b200 -lyume_pdq_cuda
#include <cstdio>
#include <cstring>
#include <vector>

#include "yume_pdq_cuda.hpp"

int main()
{
    try {
        std::vector<float> frame(512 * 512);
        for (int y = 0; y < 512; y++)
            for (int x = 0; x < 512; x++)
                frame[y * 512 + x] = float((x * 3 + y * 5) % 256);
        ypdq::CudaKernel kernel = ypdq::smart_kernel(0);
        std::vector<uint8_t> hashes(8 * 32);
        float quality[8], thr[8];
        kernel.hash_batch(frame.data(), 1, hashes.data(), quality, thr);
        std::printf("kernel: %s\nquality %.3f threshold %.4f hash ", kernel.ident(), quality[0], thr[0]);
        for (int i = 0; i < 32; i++)
            std::printf("%02x", hashes[i]);
        std::printf("\n");

        ypdq::CudaVectorDatabase db = ypdq::CudaVectorDatabase::synthetic(0, 1 << 20, 0, 0x79756d652d706471ull);
        uint8_t planted[32];
        std::memcpy(planted, hashes.data() + 32 * 5, 32);  // ROTATED_270 of the frame
        planted[0] ^= 0x0f;                                 // 4 bits away
        yume_pdq_cuda_db_write_rows(db.raw(), 123456, planted, 1);
        ypdq::CudaMatcher matcher(db, 31);
        auto found = matcher.find<std::pair<size_t, size_t>>(hashes.data(), 8, [](size_t row, size_t needle) {
            return std::optional<std::pair<size_t, size_t>>({row, needle});
        });
        if (!found || found->first != 123456 || found->second != 5) {
            std::printf("FAILED to find the planted row\n");
            return 1;
        }
        std::printf("found at row %zu, needle %zu (%s)\n", found->first, found->second, ypdq::CudaMatcher::ident());

        // the same frame as raw 8-bit luma (what `yume-pdq pipe` ingests): conversion fused into the hasher, and
        // hash + dihedrals + match in one call
        std::vector<uint8_t> luma(512 * 512);
        for (size_t i = 0; i < luma.size(); i++)
            luma[i] = uint8_t(frame[i]);
        std::vector<uint8_t> hashes8(8 * 32);
        kernel.hash_batch_pixels(luma.data(), YUME_PDQ_CUDA_PIXEL_LUMA8, 1, hashes8.data());
        if (hashes8 != hashes) {
            std::printf("FAILED: luma8 and f32 hashes differ\n");
            return 1;
        }
        auto matches = matcher.screen(luma.data(), YUME_PDQ_CUDA_PIXEL_LUMA8, 1);
        if (matches.size() != 1 || matches[0].row != 123456 || matches[0].needle != 5 || matches[0].dist != 4) {
            std::printf("FAILED: screen() returned %zu matches\n", matches.size());
            return 1;
        }
        std::printf("screen: row %llu needle %u dist %u\n", (unsigned long long)matches[0].row, matches[0].needle, matches[0].dist);
        return 0;
    } catch (const ypdq::Error &e) {
        std::printf("error: %s\n", e.what());
        return e.code == YUME_PDQ_CUDA_E_NO_DEVICE ? 2 : 1;
    }
}
