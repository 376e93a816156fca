/* A non-Python consumer of the C ABI: compiled as C (not C++) against include/wignerd_b200.h and linked against
 * libwignerd_b200.so.  Taking the address of every declared entry point makes the link fail if the header and the
 * library disagree on a name; the calls below are the ones that need no GPU.  Built and run by tests/test_host.py. */
#include <stdio.h>
#include <string.h>
#include "wignerd_b200.h"

typedef void (*fn)(void);

int main(void)
{
    fn table[] = {
        (fn)wd_create, (fn)wd_destroy, (fn)wd_last_error, (fn)wd_abi_version, (fn)wd_set_stream, (fn)wd_synchronize,
        (fn)wd_launch_count, (fn)wd_cache_clear, (fn)wd_prepare, (fn)wd_eigvecs, (fn)wd_d_batch, (fn)wd_D_batch, (fn)wd_d_matrix,
        (fn)wd_D_matrix, (fn)wd_d_multi, (fn)wd_djmn_batch, (fn)wd_Djmn_batch, (fn)wd_set_variant, (fn)wd_fp64_peak,
        (fn)wd_k2_shared_bytes, (fn)wd_create_multi, (fn)wd_multi_size, (fn)wd_multi_child, (fn)wd_shard_bounds, (fn)wd_gather,
        (fn)wd_tables_save, (fn)wd_tables_load, (fn)wd_set_slab_bytes, (fn)wd_rotate_sh,
    };
    int64_t lo = -1, hi = -1;
    unsigned i, n = 0;
    for (i = 0; i < sizeof table / sizeof table[0]; ++i) n += table[i] != 0;
    if (wd_abi_version() != 1) return 1;
    if (wd_shard_bounds(100000, 3, 8, 16, &lo, &hi) != WD_OK || lo != 37520 || hi != 50016) return 2;
    if (wd_shard_bounds(10, 0, 0, 16, &lo, &hi) != WD_ERR_ASSERT || strlen(wd_last_error()) == 0) return 3;
    if (wd_k2_shared_bytes(200, 3) <= 0 || wd_k2_shared_bytes(-1, 0) != -1) return 4;
    if (wd_create(0, 0) != WD_ERR_ASSERT) return 5;
    if (wd_d_batch(0, 2, 0, 1, 0, WD_MEM_HOST) != WD_ERR_ASSERT) return 6;        /* null handle: status code, no crash */
    if (wd_rotate_sh(0, 4, 0, 0, 0, 0, 0, 1, 0, WD_MEM_HOST) != WD_ERR_ASSERT) return 7;
    printf("c abi ok: %u entry points\n", n);
    return 0;
}
