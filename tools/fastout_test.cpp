// fastout_test: the row writer of `mash dist` / `mash triangle` (mash_b200/host/fastout.hpp) against plain `cout << ... << endl`.
//   fastout_test stream|fast <threads> <rows> <cols>  -> prints the same synthetic pair grid either way; tests/test_host_shim.py diffs the outputs.
#include "fastout.hpp"
#include <cmath>
#include <random>
using namespace std;
int main(int argc, char **argv)
{
    const string mode = argv[1];
    const int threads = atoi(argv[2]);
    const uint64_t rows = strtoull(argv[3], 0, 10), cols = strtoull(argv[4], 0, 10);
    mt19937_64 rng(rows * 1315423911u + cols);
    vector<double> d(rows * cols), p(rows * cols);
    vector<uint32_t> a(rows * cols);
    const double special[] = {0., 1., 0.5, 1e-5, 9.9999995e-5, 0.0001, 0.000123456789, 0.1234565, 0.1234575, 0.99999949, 0.9999995, 123456.7, 1234567.0, 1e-300,
                              4.9e-324, 2.2250738585072014e-308, 1e22, 0.30000000000000004, 2.5e-05, 6.25e-10, 1.0 / 3, 2.0 / 3, 1e-10, 5e-324, 1.7976931348623157e308};
    for (size_t i = 0; i < d.size(); i++) {
        const uint64_t x = rng();
        d[i] = (x % 7 == 0) ? special[(x >> 8) % (sizeof special / sizeof *special)] : (x % 7 == 1 ? (double)((x >> 8) % 1001) / 1000. : -log(2. * ((x >> 8) % 1000 + 1) / 1001. / (1. + ((x >> 8) % 1000 + 1) / 1001.)) / 21.);
        p[i] = (x % 5 == 0) ? special[(x >> 16) % (sizeof special / sizeof *special)] : exp(-(double)((x >> 20) % 700) * ((x >> 40) % 100) / 50.);
        a[i] = (uint32_t)(x >> 32) % 100001;
    }
    vector<string> names(max(rows, cols));
    for (size_t i = 0; i < names.size(); i++) names[i] = "genome_" + to_string(i * 7919 % 100003) + (i % 3 ? ".fna" : "");
    cout << "#header\t" << rows << endl;
    if (mode == "stream") {
        for (uint64_t i = 0; i < rows; i++) {
            for (uint64_t j = 0; j < cols; j++) {
                const size_t k = i * cols + j;
                if (a[k] % 4 == 0) continue;
                cout << names[j] << '\t' << names[i] << '\t' << d[k] << '\t' << p[k] << '\t' << a[k] << '/' << 1000 << endl;
            }
            cout << names[i];
            for (uint64_t j = 0; j < cols; j++) cout << '\t' << d[i * cols + j];
            cout << endl;
        }
    } else {
        mashhost::writeRows(rows, threads, cols, [&](uint64_t i, mashhost::OutBuf &o) {
            for (uint64_t j = 0; j < cols; j++) {
                const size_t k = i * cols + j;
                if (a[k] % 4 == 0) continue;
                o.str(names[j]); o.ch('\t'); o.str(names[i]); o.ch('\t'); o.dbl(d[k]); o.ch('\t'); o.dbl(p[k]); o.ch('\t'); o.u64(a[k]); o.ch('/'); o.u64(1000); o.ch('\n');
            }
            o.str(names[i]);
            for (uint64_t j = 0; j < cols; j++) { o.ch('\t'); o.dbl(d[i * cols + j]); }
            o.ch('\n');
        });
    }
    cout << "#trailer" << endl;
    return 0;
}
