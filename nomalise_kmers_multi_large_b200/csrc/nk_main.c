/* normalise_kmers_multi_large_b200 -- drop-in command line for the reference binary (C:2223-2455) */
#include "../../include/nk_b200.h"

int main(int argc, char **argv)
{
    return nk_main(argc, argv);
}
