// FASTQuick_b200: the `align` sub-command of FASTQuick (src/FASTQuick.cpp:654-672) on the B200 engine.
#include <cstdio>
#include <cstring>
#include "FastQuickB200.h"

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "align") == 0) return fqb200::runAlign(argc - 1, argv + 1);
    fprintf(stderr, "Usage: FASTQuick_b200 align --index_prefix P (--fq_list L | --fastq_1 A --fastq_2 B) --out_prefix O [--t N] [--q 15] ...\n"
                    "       (index / pop+con stay with the reference FASTQuick binary)\n");
    return 1;
}
