// CPU-only check of signal::Block in host/sdr.hpp (adapters/block.rs:106-207): prints what each reader of a tee'd
// Block sees for the schedule given on the command line; tests/test_signal_host.py compares the lines with the
// per-sample restatement in tests/pyref.py.  Needs no GPU: from_iter + Block only call host-pure entry points.
//   usage: block_tee_test <n_samples> <rate> <size> <dedup> (<reader> <count> | c <src> )...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "../../unnamed-rust-sdr_b200/host/sdr.hpp"

int main(int argc, char **argv) {
    using namespace sdr;
    if (argc < 5) return 2;
    const size_t n = (size_t)atol(argv[1]);
    const float rate = (float)atof(argv[2]), size = (float)atof(argv[3]);
    const bool dedup = atoi(argv[4]) != 0;
    std::vector<float> x(n);
    for (size_t i = 0; i < n; ++i) x[i] = (float)(i + 1);
    std::map<int, std::shared_ptr<signal::Block<float>>> readers;
    readers[0] = signal::block(signal::from_iter(rate, x), size, dedup);
    std::map<int, std::vector<float>> seen;
    for (int a = 5; a + 1 < argc; a += 2) {
        if (!strcmp(argv[a], "c")) {
            const int src = atoi(argv[a + 1]);
            const int id = (int)readers.size();
            readers[id] = readers[src]->clone();
            continue;
        }
        const int rid = atoi(argv[a]);
        size_t left = (size_t)atol(argv[a + 1]);
        while (left > 0) {
            std::vector<float> b;
            const size_t k = readers[rid]->next_block(left, b);
            if (!k) break;
            seen[rid].insert(seen[rid].end(), b.begin(), b.end());
            left -= k;
        }
    }
    for (auto &kv : seen) {
        printf("%d:", kv.first);
        for (float v : kv.second) printf(" %d", (int)v);
        printf("\n");
    }
    return 0;
}
