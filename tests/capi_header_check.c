/* include/b200ann.h must be a valid plain-C header (the boundary is a C ABI): compiled with gcc -std=c99 -pedantic. */
#include "../include/b200ann.h"

int header_is_plain_c(void) {
    ann_config cfg = {ANN_METRIC_COSINE, 200, 0, 0, ANN_FLAG_NO_SHADOW};
    ann_index *ix = 0;
    int (*create)(const ann_config *, ann_index **) = ann_create;
    int (*query)(ann_index *, const float *, int32_t, int32_t, int32_t, int64_t *, float *, int32_t *) = ann_query_batch;
    (void)cfg; (void)ix; (void)create; (void)query;
    return ANN_OK == 0 && ANN_ERR_CANDIDATE_OVERFLOW == -8;
}
