// ref_kseq_dump.cpp -- test infrastructure: dumps what the REFERENCE's own parser (kseq.h, compiled in place from
// $(REF); nothing is copied) delivers for a FASTA/FASTQ(.gz) file, the way the reference instantiates it
// (Sketch.cpp:21 `KSEQ_INIT(gzFile, gzread)`, kseq_read loop as in Sketch.cpp:1217-1270).  Used by
// tests/test_host_fastx_vs_kseq.py to pin the product's reader (mash_b200/host/fastx.hpp) record by record.
// Output: per record "R <name_len> <comment_len> <seq_len>\n" followed by the three strings, each newline-terminated;
// last line "E <return code of the final kseq_read>" (-1 end of file, -2 truncated quality).
#include <zlib.h>
#include <stdio.h>
#include "kseq.h"
KSEQ_INIT(gzFile, gzread)

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    gzFile fp = gzopen(argv[1], "r");
    if (!fp) return 3;
    kseq_t *ks = kseq_init(fp);
    int l;
    while ((l = kseq_read(ks)) >= 0) {
        printf("R %zu %zu %zu\n", (size_t)ks->name.l, (size_t)ks->comment.l, (size_t)ks->seq.l);
        fwrite(ks->name.s ? ks->name.s : "", 1, ks->name.l, stdout); fputc('\n', stdout);
        fwrite(ks->comment.s ? ks->comment.s : "", 1, ks->comment.l, stdout); fputc('\n', stdout);
        fwrite(ks->seq.s ? ks->seq.s : "", 1, ks->seq.l, stdout); fputc('\n', stdout);
    }
    printf("E %d\n", l);
    kseq_destroy(ks);
    gzclose(fp);
    return 0;
}
