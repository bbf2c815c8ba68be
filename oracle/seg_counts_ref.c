/* CPU oracle (C restatement) of the mIoU / accuracy counts -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates metrices/mIoU.py:21-29 and metrices/Accuracy.py:19-20 of the reference as one pass of exact
 * integer counting for the reference dtypes: pred int64 (argmax output,
 * train_or_resume.py:477 / benchmark.py:69), target uint8 (ignore = 255), mask bool.
 * Only tests/ and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * out[0..NC)      area_pred[c]   = #{mask & pred == c}
 * out[NC..2NC)    area_inter[c]  = #{mask & pred == c & target == c}
 * out[2NC..3NC)   area_target[c] = #{target == c}            (not masked)
 * out[3NC]        correct        = #{mask & pred == target}  (raw values)
 * out[3NC+1]      valid          = #{mask}
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void seg_counts_ref(const int64_t *pred, const uint8_t *target, const uint8_t *mask, int64_t npix,
                    int num_classes, int64_t *out, int threads)
{
    const int nc = num_classes;
    const int len = 3 * nc + 2;
    if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        int64_t *loc = (int64_t *)calloc((size_t)len, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < npix; ++i) {
            const int64_t p = pred[i];
            const int64_t t = (int64_t)target[i];
            const int m = mask[i] != 0;
            if (t < nc) loc[2 * nc + t]++;
            if (m) {
                loc[3 * nc + 1]++;
                if (p == t) loc[3 * nc]++;
                if (p >= 0 && p < nc) {
                    loc[p]++;
                    if (p == t) loc[nc + p]++;
                }
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int j = 0; j < len; ++j) out[j] += loc[j];
        free(loc);
    }
}
