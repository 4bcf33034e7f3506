/* include/psacb200.h must be a valid C header (the boundary is a C ABI): compiled with gcc -std=c99 -pedantic by
 * tests/test_abi.py.  Calls only entry points that need no GPU. */
#include <stdio.h>
#include <string.h>

#include "psacb200.h"

int main(void) {
    uint64_t start = 0, size = 0;
    psacb200_blk_dist(10, 3, 1, &start, &size); /* mxx::blk_dist: blocks of 4, 3, 3 */
    if (start != 4 || size != 3) return 1;
    psacb200_stats st;
    memset(&st, 0, sizeof st);
    if (psacb200_get_stats(NULL, &st) != PSACB200_ERR_ARG) return 2;
    if (strstr(psacb200_last_error(), "null") == NULL) return 3;
    if (psacb200_launch_count(NULL) != 0) return 4;
    psacb200_destroy(NULL);
    puts("header ok");
    return 0;
}
