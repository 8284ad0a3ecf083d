/* test-only: the device inflate routine (csrc/gf_inflate.cuh) compiled for the host, as a shared library the CPU tests call
 * with zlib-made raw DEFLATE streams.  g++ -O2 -shared -fPIC tests/c_driver/inflate_host.cpp -o tests/c_driver/libgf_inflate_host.so */
#include "../../genefuserust_b200/csrc/gf_inflate.cuh"

extern "C" int gf_test_inflate(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    static thread_local gfinf::Tables T;
    return gfinf::inflate_member(in, in_len, out, out_len, T);
}
extern "C" uint32_t gf_test_crc32(const uint8_t* p, uint32_t n) {
    static uint32_t t[4 * 256];
    static bool made = false;
    if (!made) { gfinf::crc_tables(t, 0, 1); gfinf::crc_tables_rest(t, 0, 1); made = true; }
    return gfinf::crc32_of(p, n, t);
}
/* CRC-32 of `n` bytes from the CRC-32s of `parts` slices, the way the device combines the 32 lanes' partial values */
extern "C" uint32_t gf_test_crc32_sliced(const uint8_t* p, uint32_t n, uint32_t parts) {
    static uint32_t t[4 * 256], x2n[32];
    static bool made = false;
    if (!made) { gfinf::crc_tables(t, 0, 1); gfinf::crc_tables_rest(t, 0, 1); gfinf::crc_x2n_table(x2n); made = true; }
    const uint32_t per = (n + parts - 1) / parts;
    uint32_t total = 0;
    for (uint32_t l = 0; l < parts; l++) {
        const uint32_t a = l * per < n ? l * per : n, b = (l + 1) * per < n ? (l + 1) * per : n;
        const uint32_t c = gfinf::crc32_of(p + a, b - a, t);
        total ^= b > a ? gfinf::crc_multmodp(gfinf::crc_x8n(n - b, x2n), c) : 0u;
    }
    return total;
}
