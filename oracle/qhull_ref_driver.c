/*
 * qhull_ref_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A C entry point around the reference's OWN vendored Qhull (spatial/qhull_src/src/, Qhull 2019.1 "_r", plus the
 * reference's spatial/qhull_misc.c), compiled from the sources where they lie under /root/reference by
 * oracle/Makefile into oracle/_ref/libqhull_ref.so.  No reference source is copied into this repository; this
 * file is the only code here and it restates, in C, what the reference's Cython wrapper does around Qhull for
 * `Delaunay(points)` (the wrapper itself, spatial/qhull.pyx, cannot be built: Cython 0.29 / numpy.distutils):
 *
 *   spatial/qhull.pyx:1867-1885  Delaunay.__init__: options "Qbb Qc Qz Q12" (ndim < 5) + required "Qt", mode "d"
 *   spatial/qhull.pyx:338-363    _Qhull.__init__:  "qhull d <options>" -> qh_zero + qh_new_qhull_scipy(..., outfile=NULL)
 *   spatial/qhull.pyx:563-568    _Qhull.triangulate: qh_triangulate
 *   spatial/qhull.pyx:573-723    get_simplex_facet_array: lower-Delaunay facets in facet_list order, vertex /
 *                                neighbour k from facet->vertices / facet->neighbors, first two swapped for
 *                                clockwise facets, -1 for neighbours that are not lower-Delaunay facets
 *
 * The option ORDER matters to nothing in Qhull (flags), but is kept as the wrapper produces it is impossible (it joins
 * a Python set); the flags themselves are what is reproduced.
 *
 * int qhull_ref_delaunay2d(const double* points, int n, int* simplices, int* neighbors, int max_facets)
 *   points [n,2] row-major (the reference passes (row, col) pixel coordinates, interp2d.py:53-55)
 *   simplices / neighbors [max_facets,3] int32 out
 *   returns the number of triangles (>= 0), -1 on a Qhull error, -2 if max_facets is too small.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qhull_misc.h"

int qhull_ref_delaunay2d(const double* points, int n, int* simplices, int* neighbors, int max_facets) {
  qhT* qh = (qhT*)malloc(sizeof(qhT));
  FILE* err = tmpfile();
  char cmd[] = "qhull d Qbb Qc Qz Q12 Qt";
  int result = -1, curlong = 0, totlong = 0;
  int* id_map = NULL;
  facetT* facet;
  if (!qh || !err) goto done;
  qh_zero(qh, err);
  /* ismalloc = 0: Qhull reads the caller's array in place, exactly as the wrapper passes points.data */
  if (qh_new_qhull_scipy(qh, 2, n, (coordT*)points, 0, cmd, NULL, err, NULL) != 0) goto cleanup;
  qh_triangulate(qh);

  {
    int nid = (int)qh->facet_id, j = 0, i;
    const int facet_ndim = 3;
    id_map = (int*)malloc(sizeof(int) * (size_t)(nid > 0 ? nid : 1));
    if (!id_map) goto cleanup;
    for (i = 0; i < nid; ++i) id_map[i] = -1;
    for (facet = qh->facet_list; facet && facet->next; facet = facet->next) {
      if (facet->upperdelaunay == qh->UPPERdelaunay) {
        if (!facet->simplicial && (qh_setsize(qh, facet->vertices) != facet_ndim ||
                                   qh_setsize(qh, facet->neighbors) != facet_ndim))
          goto cleanup; /* "non-simplical facet encountered" in the wrapper */
        id_map[facet->id] = j++;
      }
    }
    if (j > max_facets) { result = -2; goto cleanup; }
    j = 0;
    for (facet = qh->facet_list; facet && facet->next; facet = facet->next) {
      int lower = 0;
      if (facet->upperdelaunay != qh->UPPERdelaunay) continue;
      if (facet->toporient == qh_ORIENTclock) {
        for (i = 0; i < 2; ++i) {
          const int swapped = 1 ^ i;
          vertexT* v = (vertexT*)facet->vertices->e[i].p;
          facetT* nb = (facetT*)facet->neighbors->e[i].p;
          simplices[3 * j + swapped] = qh_pointid(qh, v->point);
          neighbors[3 * j + swapped] = id_map[nb->id];
        }
        lower = 2;
      }
      for (i = lower; i < facet_ndim; ++i) {
        vertexT* v = (vertexT*)facet->vertices->e[i].p;
        facetT* nb = (facetT*)facet->neighbors->e[i].p;
        simplices[3 * j + i] = qh_pointid(qh, v->point);
        neighbors[3 * j + i] = id_map[nb->id];
      }
      ++j;
    }
    result = j;
  }

cleanup:
  qh_freeqhull(qh, qh_ALL);
  qh_memfreeshort(qh, &curlong, &totlong);
done:
  free(id_map);
  if (err) fclose(err);
  free(qh);
  return result;
}

/*
 * int qhull_ref_delaunay2d_probe(points, n, simplices, owner, vertex_id, max_facets)
 *   The same Qhull run, reporting what decides the triangulation INSIDE co-circular cells: `Qt` (qh_triangulate,
 *   spatial/qhull_src/src/poly2_r.c) splits every merged non-simplicial facet into a fan around the facet's first
 *   vertex -- the one with the largest vertex id, i.e. the vertex Qhull's incremental hull inserted last.
 *     simplices [T,3]  point ids of every lower-Delaunay triangle, facet_list order (no orientation swap)
 *     owner     [T]    -1 for an ordinary simplicial facet, else the id of the merged facet the triangle was cut from
 *     vertex_id [n]    Qhull's vertex id of every input point (-1: not a vertex, e.g. a coplanar duplicate)
 *   returns T, -1 on a Qhull error, -2 if max_facets is too small.
 */
int qhull_ref_delaunay2d_probe(const double* points, int n, int* simplices, int* owner, int* vertex_id, int max_facets) {
  qhT* qh = (qhT*)malloc(sizeof(qhT));
  FILE* err = tmpfile();
  char cmd[] = "qhull d Qbb Qc Qz Q12 Qt";
  int result = -1, curlong = 0, totlong = 0, i, j = 0;
  facetT* facet;
  vertexT* vertex;
  if (!qh || !err) goto done;
  qh_zero(qh, err);
  if (qh_new_qhull_scipy(qh, 2, n, (coordT*)points, 0, cmd, NULL, err, NULL) != 0) goto cleanup;
  qh_triangulate(qh);
  for (i = 0; i < n; ++i) vertex_id[i] = -1;
  for (vertex = qh->vertex_list; vertex && vertex->next; vertex = vertex->next) {
    const int pid = qh_pointid(qh, vertex->point);
    if (pid >= 0 && pid < n) vertex_id[pid] = (int)vertex->id;
  }
  for (facet = qh->facet_list; facet && facet->next; facet = facet->next) {
    if (facet->upperdelaunay != qh->UPPERdelaunay) continue;
    if (j >= max_facets) { result = -2; goto cleanup; }
    for (i = 0; i < 3; ++i) simplices[3 * j + i] = qh_pointid(qh, ((vertexT*)facet->vertices->e[i].p)->point);
    owner[j] = facet->tricoplanar ? (facet->f.triowner ? (int)facet->f.triowner->id : -2) : -1;
    ++j;
  }
  result = j;
cleanup:
  qh_freeqhull(qh, qh_ALL);
  qh_memfreeshort(qh, &curlong, &totlong);
done:
  if (err) fclose(err);
  free(qh);
  return result;
}

const char* qhull_ref_version(void) { return qh_version; }
