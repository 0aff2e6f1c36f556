/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C restatement of the reference's all-pairs hop-distance
 * preprocessing, pre_process_datasets.py:104-142:
 *   hop   = scipy.sparse.csgraph.dijkstra(adj)   (:110,:129)  directed, unit weights  == BFS levels
 *   node_distances       = 1/(1+hop), inf -> 0    (:112-114)   float32 arithmetic
 *   normalization_matrix = count of equal entries in the same row (:117-121) == BFS level sizes
 * scipy (1.18.1 here; unpinned upstream) is a third-party dependency absent from /root/reference; for
 * unit weights Dijkstra's result is the BFS level, which is what is restated here. Pinned against
 * scipy + the real pre_process() by oracle/make_golden.py -> tests/golden/preprocess_*.npz.
 * Input must be a simple directed graph (the reference's COO->LIL conversion SUMS duplicate edges
 * into weight 2, pre_process_datasets.py:109; not emulated).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int gnan_oracle_apsp_rows(int32_t n, int64_t n_edges, const int64_t *src, const int64_t *dst, int32_t n_sources, int32_t *hop);

/* hop: int32 [n*n], -1 = unreachable. returns 0 ok, 1 alloc failure, 2 bad edge. */
int gnan_oracle_apsp(int32_t n, int64_t n_edges, const int64_t *src, const int64_t *dst, int32_t *hop)
{
    return gnan_oracle_apsp_rows(n, n_edges, src, dst, n, hop);
}

/* BFS from sources 0..n_sources-1 only: hop is int32 [n_sources*n]. */
int gnan_oracle_apsp_rows(int32_t n, int64_t n_edges, const int64_t *src, const int64_t *dst, int32_t n_sources, int32_t *hop)
{
    int64_t *rowptr = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int32_t *col = (int32_t *)malloc((size_t)(n_edges > 0 ? n_edges : 1) * sizeof(int32_t));
    int32_t *queue = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (!rowptr || !col || !queue) { free(rowptr); free(col); free(queue); return 1; }
    for (int64_t e = 0; e < n_edges; ++e) {
        if (src[e] < 0 || src[e] >= n || dst[e] < 0 || dst[e] >= n) { free(rowptr); free(col); free(queue); return 2; }
        rowptr[src[e] + 1]++;
    }
    for (int32_t v = 0; v < n; ++v) rowptr[v + 1] += rowptr[v];
    int64_t *fill = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    if (!fill) { free(rowptr); free(col); free(queue); return 1; }
    memcpy(fill, rowptr, (size_t)n * sizeof(int64_t));
    for (int64_t e = 0; e < n_edges; ++e) col[fill[src[e]]++] = (int32_t)dst[e];
    free(fill);
    for (int32_t s = 0; s < n_sources && s < n; ++s) {
        int32_t *row = hop + (int64_t)s * n;
        for (int32_t v = 0; v < n; ++v) row[v] = -1;
        int32_t head = 0, tail = 0;
        row[s] = 0; queue[tail++] = s;
        while (head < tail) {
            int32_t u = queue[head++];
            for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
                int32_t w = col[e];
                if (row[w] < 0) { row[w] = row[u] + 1; queue[tail++] = w; }
            }
        }
    }
    free(rowptr); free(col); free(queue);
    return 0;
}

/* cnt: int32 [n_rows*nbins]; bin d = #{j: hop[i,j]==d} for d<nbins-1, last bin = unreachable. returns max finite hop, or -2 if a hop >= nbins-1. */
int gnan_oracle_level_counts(int32_t n_rows, int32_t n, const int32_t *hop, int32_t nbins, int32_t *cnt)
{
    int maxd = 0;
    memset(cnt, 0, (size_t)n_rows * nbins * sizeof(int32_t));
    for (int32_t i = 0; i < n_rows; ++i)
        for (int32_t j = 0; j < n; ++j) {
            int32_t h = hop[(int64_t)i * n + j];
            if (h < 0) { cnt[(int64_t)i * nbins + nbins - 1]++; continue; }
            if (h >= nbins - 1) return -2;
            if (h > maxd) maxd = h;
            cnt[(int64_t)i * nbins + h]++;
        }
    return maxd;
}

/* reference-format outputs (float32 [n_rows*n] each). */
void gnan_oracle_reference_format(int32_t n_rows, int32_t n, const int32_t *hop, int32_t nbins, const int32_t *cnt,
                                  float *node_distances, float *normalization_matrix)
{
    for (int32_t i = 0; i < n_rows; ++i)
        for (int32_t j = 0; j < n; ++j) {
            int64_t o = (int64_t)i * n + j;
            int32_t h = hop[o];
            node_distances[o] = h < 0 ? 0.0f : 1.0f / ((float)h + 1.0f);
            normalization_matrix[o] = (float)cnt[(int64_t)i * nbins + (h < 0 ? nbins - 1 : h)];
        }
}
