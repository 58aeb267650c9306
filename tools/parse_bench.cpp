// parse_bench: how fast do T threads turn FASTA files into sequence strings (the CLI's parse stage, no GPU)?
#include "fastx.hpp"
#include <atomic>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sys/stat.h>
#include <thread>
using namespace std;
int main(int argc, char **argv)
{
    int T = atoi(argv[1]);
    int keep = argc > 3 ? atoi(argv[3]) : 0;
    vector<string> files; { ifstream in(argv[2]); string l; while (getline(in, l)) files.push_back(l); }
    atomic<size_t> next(0); atomic<uint64_t> bytes(0);
    auto t0 = chrono::steady_clock::now();
    vector<thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&]() {
        vector<vector<mashhost::SeqBuffer>> held;
        for (;;) {
            size_t i = next++;
            if (i >= files.size()) break;
            gzFile fp = mashhost::FastxReader::openPath(files[i]);
            mashhost::FastxReader r(fp);
            struct stat st; if (stat(files[i].c_str(), &st) == 0) r.setSizeHint(st.st_size);
            vector<mashhost::SeqBuffer> seqs; int l; uint64_t b = 0;
            while ((l = r.read()) >= 0) { b += r.seq.size(); seqs.push_back(std::move(r.seq)); }
            gzclose(fp);
            bytes += b;
            if (keep) { held.push_back(std::move(seqs)); if ((int)held.size() > keep) held.erase(held.begin()); }
        }
    });
    for (auto &x : th) x.join();
    double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    printf("T=%d files=%zu bases=%.2f GB  %.3f s  %.2f GB/s\n", T, files.size(), bytes / 1e9, s, bytes / 1e9 / s);
}
