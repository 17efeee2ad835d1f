// TEST INFRASTRUCTURE: the string-interning kernels (intern_kernel.cu) compiled for the host and run under cuda_emu.h in
// the order of launch_intern: equal strings must get equal ids, ids must number the distinct strings in order of first
// occurrence (what iadh_intern_strings promises, include/iadhore_b200.h) -- compared with a std::unordered_map.
//
// input (little endian): i64 n; i64 off[n + 1]; char blob[off[n]]
#define IADH_HOST_EMU 1
#define EMU_MAX_THREADS 1024
#define EMU_STACK_BYTES (256 << 10)
#include <climits>
#include <string>
#include <unordered_map>

#include "cuda_emu.h"

#include "../../i-adhore_b200/csrc/intern_kernel.cu"

#include "emu_runtime.inc"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <strings.bin>\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) return 2;
    std::vector<int64_t> off(n + 1);
    if (fread(off.data(), 8, n + 1, f) != (size_t)(n + 1)) return 2;
    std::vector<char> blob((size_t)off[n] + 1);
    if (off[n] && fread(blob.data(), 1, (size_t)off[n], f) != (size_t)off[n]) return 2;
    fclose(f);
    uint32_t table_size = 16;
    while ((int64_t)table_size < 2 * n) table_size <<= 1;
    std::vector<int32_t> table(table_size, -1), rep(n + 1, -7), cmin(n + 1, INT_MAX), flag_scan(n + 1, -7), block_sum(n / 1024 + 2, 0), ids(n + 1, -7);
    int32_t total = -1;
    if (n > 0) {
        const unsigned g256 = (unsigned)((n + 255) / 256), g1024 = (unsigned)((n + 1023) / 1024);
        emu::launch(g256, 256, [&]() { intern_insert_kernel(blob.data(), off.data(), n, table.data(), table_size - 1, rep.data()); });
        emu::launch(g256, 256, [&]() { intern_min_kernel(rep.data(), n, cmin.data()); });
        emu::launch(g1024, 1024, [&]() { intern_flag_kernel(rep.data(), cmin.data(), n, flag_scan.data(), block_sum.data()); });
        emu::launch(1, 1024, [&]() { intern_block_scan_kernel(block_sum.data(), (int)g1024, &total); });
        emu::launch(g256, 256, [&]() { intern_ids_kernel(rep.data(), cmin.data(), flag_scan.data(), block_sum.data(), n, ids.data()); });
    } else
        total = 0;
    std::unordered_map<std::string, int32_t> seen;
    long long bad = 0;
    for (int64_t i = 0; i < n; i++) {
        const std::string s(blob.data() + off[i], (size_t)(off[i + 1] - off[i]));
        const int32_t want = seen.emplace(s, (int32_t)seen.size()).first->second;
        bad += ids[i] != want;
    }
    bad += total != (int32_t)seen.size();
    printf("{\"strings\": %lld, \"distinct\": %d, \"mismatching\": %lld}\n", (long long)n, total, bad);
    return bad ? 1 : 0;
}
