// Host driver for host/tiff_io.hpp, used by tests/test_tiff_cpu.py:
//   tiff_io_host_test write <out.tiff> <width> <height> <spp> <raw u16 file> [lzw]
//   tiff_io_host_test read  <in.tiff> <raw u16 out>      (prints "width height spp")
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../opticalimageprocessor_b200/host/tiff_io.hpp"

int main(int argc, char **argv)
{
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "write" && (argc == 7 || argc == 8)) {
            const long w = atol(argv[3]), h = atol(argv[4]);
            const int spp = atoi(argv[5]);
            std::vector<uint16_t> px((size_t)w * h * spp);
            FILE *f = fopen(argv[6], "rb");
            if (!f || fread(px.data(), 2, px.size(), f) != px.size()) return 3;
            fclose(f);
            oiptiff::write_u16(argv[2], px.data(), w, h, spp, spp == 1 ? 1 : 2, argc == 8 ? oiptiff::COMPRESS_LZW : oiptiff::COMPRESS_NONE);
            return 0;
        }
        if (mode == "read" && argc == 4) {
            const oiptiff::Info I = oiptiff::read_info(argv[2]);
            std::vector<uint16_t> px((size_t)I.width * I.height * I.spp);
            oiptiff::read_u16(argv[2], I, px.data());
            FILE *f = fopen(argv[3], "wb");
            if (!f || fwrite(px.data(), 2, px.size(), f) != px.size()) return 3;
            fclose(f);
            printf("%lld %lld %d\n", (long long)I.width, (long long)I.height, I.spp);
            return 0;
        }
        fprintf(stderr, "usage error\n");
        return 2;
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
