// ThreadSanitizer driver for the concurrent parts of csrc/host_text.cpp (block formatter / ordered writer, concurrent sample
// parsing, concurrent gene lookup, BED12 parsed in pieces).  Not part of the library or of pytest:
//   g++ -std=c++17 -O1 -g -fsanitize=thread -Iinclude tests/tools/tsan_host_text.cpp spliser_b200/csrc/host_text.cpp -o /tmp/tsan_ht -lpthread && /tmp/tsan_ht
// Exit code 0 and no "WARNING: ThreadSanitizer" on stderr = clean.  Output correctness is pinned elsewhere (tests/test_cli_*.py);
// here the same table is written with 1 worker and with all workers and the two files must be identical.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "spliser_b200.h"

static std::string slurp(const std::string& p) { std::ifstream f(p, std::ios::binary); std::stringstream s; s << f.rdbuf(); return s.str(); }

int main() {
    const int64_t n = 200000;
    std::vector<int32_t> chrom(n), pos(n), ppos, cpos, line_strand(n), site_gene(n);
    std::vector<int64_t> first(n), alpha(n), b1(n), b2(n), bc(n), poff(n + 1, 0), pcnt, coff(n + 1, 0);
    std::vector<double> bw(n), sse(n);
    uint64_t x = 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (int64_t i = 0; i < n; ++i) {
        chrom[i] = (int32_t)(i * 3 / n); pos[i] = (int32_t)(100 + i * 7); first[i] = i; line_strand[i] = (int32_t)(rnd() % 3);
        alpha[i] = (int64_t)(rnd() % 500); b1[i] = (int64_t)(rnd() % 90); b2[i] = (int64_t)(rnd() % 40); bc[i] = (int64_t)(rnd() % 9);
        bw[i] = (double)(rnd() % 100000) / 977.0; sse[i] = (double)(rnd() % 1000) / 999.0; site_gene[i] = (int32_t)(rnd() % 4) - 1;
        for (int k = (int)(rnd() % 3); k > 0; --k) { ppos.push_back(pos[i] + 50 * k); pcnt.push_back((int64_t)(rnd() % 70)); }
        for (int k = (int)(rnd() % 3); k > 0; --k) cpos.push_back(pos[i] + 9 * k);
        poff[i + 1] = (int64_t)ppos.size(); coff[i + 1] = (int64_t)cpos.size();
    }
    spl_site_columns t{n, chrom.data(), pos.data(), first.data(), alpha.data(), b1.data(), b2.data(), bc.data(), bw.data(), sse.data(),
                       poff.data(), ppos.data(), pcnt.data(), coff.data(), cpos.data()};
    const int64_t coffs[] = {0, 4, 8, 12}, soffs[] = {0, 1, 2, 3}, goffs[] = {0, 2, 4, 6};
    spl_strtab ctab{3, "Chr1Chr2Chr3", coffs}, stab{3, "+-?", soffs}, gtab{3, "G1G2G3", goffs};
    char err[256];
    std::string out[2];
    for (int pass = 0; pass < 2; ++pass) {
        setenv("SPLISER_HOST_THREADS", pass ? "0" : "1", 1);
        const std::string p = "/tmp/tsan_ht_" + std::to_string(pass) + ".tsv";
        if (spl_write_process_tsv(p.c_str(), &t, &ctab, &stab, line_strand.data(), &gtab, site_gene.data(), pass, err, 256) != 0) { fprintf(stderr, "write: %s\n", err); return 1; }
        if (pass == 0) spl_write_process_tsv(p.c_str(), &t, &ctab, &stab, line_strand.data(), &gtab, site_gene.data(), 1, err, 256);
        out[pass] = slurp(p);
    }
    if (out[0] != out[1] || out[0].size() < 1000000) { fprintf(stderr, "1 worker and all workers wrote different files\n"); return 1; }
    // the written table as three "samples" (one truncated, one thinned) through the concurrent parser, the merge and the combined writer
    unsetenv("SPLISER_HOST_THREADS");
    std::vector<std::string> paths;
    for (int k = 0; k < 3; ++k) {
        std::istringstream in(out[1]);
        std::string line, text;
        int64_t r = 0;
        while (std::getline(in, line)) { if (r == 0 || (k == 0) || (k == 1 && r < n / 2) || (k == 2 && r % 3)) text += line + "\n"; ++r; }
        paths.push_back("/tmp/tsan_ht_s" + std::to_string(k) + ".tsv");
        std::ofstream(paths.back(), std::ios::binary) << text;
    }
    spl_combine* cm = nullptr;
    if (spl_combine_create(&cm) != 0) return 1;
    const char* titles[] = {"a", "b", "c"};
    const char* pp[] = {paths[0].c_str(), paths[1].c_str(), paths[2].c_str()};
    if (spl_combine_add_samples(cm, 3, titles, pp, 0) != 0) { fprintf(stderr, "add: %s\n", spl_combine_last_error(cm)); return 1; }
    std::vector<int32_t> order;
    for (int64_t r = 0; r < spl_combine_n_regions(cm); ++r) order.push_back((int32_t)r);
    if (spl_combine_merge_shallow(cm, (int64_t)order.size(), order.data(), nullptr, 1, 2, 5, 0.1) != 0) { fprintf(stderr, "merge: %s\n", spl_combine_last_error(cm)); return 1; }
    for (int k = 0; k < 3; ++k) {
        const int64_t g = spl_combine_gaps(cm, k, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        std::vector<int64_t> z((size_t)g + 1, 3);
        if (spl_combine_set_recount(cm, k, g, z.data(), z.data()) != 0) return 1;
    }
    if (spl_combine_write(cm, "/tmp/tsan_ht_combined.tsv", 1) != 0) { fprintf(stderr, "cwrite: %s\n", spl_combine_last_error(cm)); return 1; }
    printf("sites %lld filled %lld combined bytes %zu\n", (long long)spl_combine_n_sites(cm), (long long)spl_combine_n_filled(cm), slurp("/tmp/tsan_ht_combined.tsv").size());
    spl_combine_destroy(cm);
    // gene lookup in concurrent blocks + BED12 in concurrent pieces
    std::vector<int32_t> gl, gr, gs, q(n), qs(n), idx(n);
    for (int k = 0; k < 5000; ++k) { gl.push_back(k * 300); gr.push_back(k * 300 + 250); gs.push_back(k & 1); }
    for (int64_t i = 0; i < n; ++i) { q[i] = (int32_t)(rnd() % 1500000); qs[i] = (int32_t)(rnd() % 3); }
    if (spl_gene_search(5000, gl.data(), gr.data(), gs.data(), n, q.data(), qs.data(), 0, 1, 1, idx.data()) != 0) return 1;
    std::string bed;
    for (int k = 0; k < 60000; ++k) bed += "Chr" + std::to_string(1 + k / 20000) + "\t" + std::to_string(k * 40) + "\t" + std::to_string(k * 40 + 300) + "\tJ\t" + std::to_string(1 + k % 50) + "\t+\t0\t0\t0\t2\t10,12\t0,90\n";
    spl_bed* b = nullptr;
    if (spl_bed_parse(bed.data(), (int64_t)bed.size(), nullptr, nullptr, 0, 0, 0, 0, &b, err, 256) != 0) { fprintf(stderr, "bed: %s\n", err); return 1; }
    printf("bed rows %lld\n", (long long)spl_bed_n_junctions(b));
    spl_bed_free(b);
    return 0;
}
