// developer tool: zlib inflate of BGZF members by N threads (what gf_fastq_stream does with .fq.gz written by bgzip / bcl2fastq).  g++ -O2 -pthread tools/bgzf_bench.cpp -lz
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cstdint>
struct Blk { const uint8_t* data; uint32_t clen, isize, crc; uint64_t out; };
static bool inflate_range(const std::vector<Blk>& b, size_t lo, size_t hi, uint8_t* dst, bool do_crc) {
    z_stream z; memset(&z, 0, sizeof(z)); if (inflateInit2(&z, -15) != Z_OK) return false; bool ok = true;
    for (size_t i = lo; i < hi && ok; i++) {
        z.next_in = (Bytef*)b[i].data; z.avail_in = b[i].clen; z.next_out = dst + b[i].out; z.avail_out = b[i].isize;
        int zr = inflate(&z, Z_FINISH);
        ok = zr == Z_STREAM_END && z.avail_out == 0 && (!do_crc || (uint32_t)crc32(0L, dst + b[i].out, b[i].isize) == b[i].crc);
        inflateReset(&z);
    }
    inflateEnd(&z); return ok;
}
int main(int argc, char** argv) {
    const size_t nrec = 400000; std::vector<uint8_t> text; text.reserve(nrec * 333);
    uint64_t x = 88172645463325252ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (size_t i = 0; i < nrec; i++) {
        char name[40]; int n = snprintf(name, sizeof name, "@SYN:12:%09zu 1:N:0:ACGT\n", i); text.insert(text.end(), name, name + n);
        for (int k = 0; k < 150; k++) text.push_back("ACGT"[rnd() & 3]);
        text.push_back('\n'); text.push_back('+'); text.push_back('\n');
        for (int k = 0; k < 150; k++) text.push_back("EEEEEEA/<"[rnd() % 9]);
        text.push_back('\n');
    }
    std::vector<uint8_t> comp; std::vector<Blk> blocks; std::vector<size_t> offs;
    for (size_t p = 0; p < text.size(); p += 65280) {
        size_t n = std::min<size_t>(65280, text.size() - p); uLongf cl = compressBound(n); std::vector<uint8_t> tmp(cl + 64);
        z_stream d; memset(&d, 0, sizeof d); deflateInit2(&d, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        d.next_in = &text[p]; d.avail_in = n; d.next_out = tmp.data(); d.avail_out = tmp.size(); deflate(&d, Z_FINISH); size_t got = tmp.size() - d.avail_out; deflateEnd(&d);
        offs.push_back(comp.size()); comp.insert(comp.end(), tmp.begin(), tmp.begin() + got);
        blocks.push_back(Blk{nullptr, (uint32_t)got, (uint32_t)n, (uint32_t)crc32(0L, &text[p], n), p});
    }
    for (size_t i = 0; i < blocks.size(); i++) blocks[i].data = comp.data() + offs[i];
    std::vector<uint8_t> out(text.size());
    printf("text %.1f MB, compressed %.1f MB, %zu blocks, zlib %s\n", text.size() / 1e6, comp.size() / 1e6, blocks.size(), zlibVersion());
    for (int crc = 0; crc < 2; crc++)
        for (int nt : {1, 2, 4, 8, 16}) {
            auto t0 = std::chrono::steady_clock::now(); std::vector<std::thread> th; size_t nb = blocks.size();
            for (int u = 0; u < nt; u++) th.emplace_back([&, u] { inflate_range(blocks, nb * u / nt, nb * (u + 1) / nt, out.data(), crc); });
            for (auto& t : th) t.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("crc %d threads %2d: %.1f ms, %.2f GB/s\n", crc, nt, s * 1e3, text.size() / s / 1e9);
        }
    // streaming single member for comparison
    return memcmp(out.data(), text.data(), text.size()) != 0;
}
