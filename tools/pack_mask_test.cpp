// pack_mask_test: pack_chunk_mask (the screen feed's host packer: 2-bit codes + invalid-position mask, AVX-512 / AVX2 / scalar paths, the
// persistent worker pool) against a byte-by-byte restatement.  Built and run by tests/test_host_pack.py; exit code 0 = identical.
#include "pack.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <sys/wait.h>
#include <unistd.h>
using namespace mashgpu;
int main()
{
    std::mt19937_64 rng(5);
    const char *alpha = "ACGTacgtN*\n\0xyz";
    int bad = 0;
    for (int trial = 0; trial < 300; trial++) {
        uint64_t len = (trial < 50) ? trial : rng() % 3000000 + 1;
        if (trial == 299) len = 40000000;
        std::vector<uint8_t> src(len + 64, 'A');
        for (uint64_t i = 0; i < len; i++) { int r = rng() % 100; src[i] = r < 90 ? alpha[rng() % 4] : alpha[rng() % 15]; }
        for (int pc = 0; pc < 2; pc++) for (int th : {1, 5}) {
            uint64_t groups = (len + 31) / 32;
            std::vector<uint64_t> codes(groups + 1, 0x1234);
            std::vector<uint32_t> inval(groups + 1, 0x5678);
            pack_chunk_mask(src.data(), len, pc, th, codes.data(), inval.data());
            for (uint64_t p = 0; p < groups * 32; p++) {
                int want_inv = 1, want_code = 0;
                if (p < len) {
                    int b = src[p]; if (!pc && b > 96 && b < 123) b -= 32;
                    want_code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
                    want_inv = want_code < 0;
                }
                int inv = (inval[p / 32] >> (p % 32)) & 1, code = (codes[p / 32] >> (2 * (p % 32))) & 3;
                if (inv != want_inv || (!want_inv && code != want_code)) { if (bad++ < 10) printf("trial %d len %llu pc %d th %d pos %llu: inv %d/%d code %d/%d\n", trial, (unsigned long long)len, pc, th, (unsigned long long)p, inv, want_inv, code, want_code); }
            }
            if (codes[groups] != 0x1234 || inval[groups] != 0x5678) { printf("overrun trial %d\n", trial); bad++; }
        }
    }
    {   // a forked child must be able to pack too (the pool's worker threads do not exist there)
        fflush(stdout);
        const pid_t pid = fork();
        if (pid == 0) {
            std::vector<uint8_t> src(3000000, 'C');
            std::vector<uint64_t> codes(src.size() / 32 + 1); std::vector<uint32_t> inval(src.size() / 32 + 1);
            pack_chunk_mask(src.data(), src.size(), 0, 5, codes.data(), inval.data());
            _exit(codes[100] == 0x5555555555555555ull && inval[100] == 0 ? 0 : 3);
        }
        int status = 0;
        waitpid(pid, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) { printf("forked child failed (status %d)\n", status); bad++; }
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
