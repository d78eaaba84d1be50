// cli_main.cpp -- the `cfrk` command: same positional interface as the reference's main()
// (src/main.cu:232-250), so test/test.sh and swift/roda.sh keep working:
//
//     cfrk <dataset.fasta> <file_out.cfrk> <k> [<nt>]            (argc == 5 parses nt)
//     cfrk <dataset.fasta> <file_out.cfrk> <k> <nt> <chunkSize>  (argc == 6 parses chunkSize only)
//
// Like the reference: fewer than 3 arguments prints the usage line (no newline) and returns 1;
// success prints nothing; a missing input exits 1.  Extensions (ignored by positional counting):
//     --all-rows   print every read, not only the last nS mod chunkSize
//     --exact      intended semantics instead of the reference's quirks (all windows, no spill,
//                  wrapped FASTA lines joined, last base kept)
//     --sparse     omit zero bins (the filter commented out at src/main.cu:51,56)
//                  k = 9..31 needs --sparse --exact: rows of "kmer_index:count " for the k-mers present
//     --device=N
//     --devices=0,1,2 | --devices=all   spread the file over several GPUs (rows stay in read order)
// The input may be gzip-compressed.
// Legacy Swift form (swift/cfrk.swf:5): `cfrk <dataset> <k> <chunkSize>` with numeric 2nd/3rd
// arguments writes the rows to stdout.
#include "cfrk_b200.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool all_digits(const char* s)
{
    if (!*s) return false;
    for (; *s; s++) if (!isdigit((unsigned char)*s)) return false;
    return true;
}

int main(int argc, char** argv)
{
    int flags = 0;
    std::vector<int> devices;
    std::vector<char*> pos;
    pos.push_back(argv[0]);
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--all-rows")) flags |= CFRK_RUN_ALL_ROWS;
        else if (!strcmp(argv[i], "--exact")) flags |= CFRK_RUN_EXACT;
        else if (!strcmp(argv[i], "--sparse")) flags |= CFRK_RUN_SPARSE;
        else if (!strncmp(argv[i], "--device=", 9)) devices.assign(1, atoi(argv[i] + 9));
        else if (!strcmp(argv[i], "--devices=all")) {
            devices.clear();
            for (int d = 0; d < cfrk_device_count(); d++) devices.push_back(d);
        }
        else if (!strncmp(argv[i], "--devices=", 10)) {
            devices.clear();
            for (const char* p = argv[i] + 10; *p;) {
                devices.push_back(atoi(p));
                while (*p && *p != ',') p++;
                if (*p == ',') p++;
            }
        }
        else pos.push_back(argv[i]);
    }
    const int pc = (int)pos.size();
    if (pc < 4) {
        printf("Usage: ./cfrk [dataset.fasta] [file_out.cfrk] [k] <number of threads: Default 12> <chunkSize: Default 8192>");
        return 1;
    }
    long chunk = 8192;
    int nt = 12, k;
    const char* out = pos[2];
    if (pc == 4 && all_digits(pos[2]) && all_digits(pos[3])) {  // swift/cfrk.swf:5
        k = atoi(pos[2]);
        chunk = atol(pos[3]);
        out = "/dev/stdout";
    } else {
        k = atoi(pos[3]);
        if (pc == 5) nt = atoi(pos[4]);
        if (pc == 6) chunk = atol(pos[5]);
    }
    if (devices.empty()) devices.push_back(0);
    int rc = cfrk_run_file_multi(pos[1], out, k, nt, chunk, flags, devices.data(), (int)devices.size());
    if (rc != CFRK_OK) {
        fprintf(stderr, "cfrk: error %d: %s\n", rc, cfrk_last_error());
        return 1;
    }
    return 0;
}
