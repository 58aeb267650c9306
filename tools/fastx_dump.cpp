// fastx_dump.cpp -- the product's FASTA/FASTQ reader (mash_b200/host/fastx.hpp) in the dump format of
// oracle/ref_kseq_dump.cpp, for the record-by-record comparison in tests/test_host_fastx_vs_kseq.py.
#include <cstdio>
#include "../mash_b200/host/fastx.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    gzFile fp = mashhost::FastxReader::openPath(argv[1]);
    if (!fp) return 3;
    int l;
    {
        mashhost::FastxReader reader(fp);
        while ((l = reader.read()) >= 0) {
            printf("R %zu %zu %zu\n", reader.name.size(), reader.comment.size(), reader.seq.size());
            fwrite(reader.name.data(), 1, reader.name.size(), stdout); fputc('\n', stdout);
            fwrite(reader.comment.data(), 1, reader.comment.size(), stdout); fputc('\n', stdout);
            fwrite(reader.seq.data(), 1, reader.seq.size(), stdout); fputc('\n', stdout);
        }
        printf("E %d\n", l);
    }
    gzclose(fp);
    return 0;
}
