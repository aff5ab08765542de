// edgelist.h -- host-side mirror of the reference's edge-list types and loaders
// (narayanan2004/GraphMat include/GMDP/utils/edgelist.h:38-78,242-334 and
// edgelist_transformation.h:431-443): same names and argument meaning, single process.
// File format (binary): int m, n, nnz header, then (int src, int dst, T val) records,
// ids 1-based.  The header's nnz is authoritative (the reference reads to EOF into a buffer
// sized from the header and overruns by one record on the shipped data files, SURVEY.md
// hazard 6); records beyond it are ignored here.
#ifndef GRAPHMAT_B200_EDGELIST_H
#define GRAPHMAT_B200_EDGELIST_H
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>

namespace GraphMat {

inline int get_global_nrank() { return 1; }
inline int get_global_myrank() { return 0; }

template <typename T>
struct edge_t {
  edge_t() {}
  edge_t(int _src, int _dst, T _val) : src(_src), dst(_dst), val(_val) {}
  int src;
  int dst;
  T val;
};

template <typename T>
struct edgelist_t {
  edge_t<T>* edges;
  int m;
  int n;
  int nnz;
  edgelist_t() : edges(nullptr), m(0), n(0), nnz(0) {}
  edgelist_t(int _m, int _n, int _nnz) : edges(nullptr), m(_m), n(_n), nnz(_nnz) {
    if (nnz > 0) edges = reinterpret_cast<edge_t<T>*>(malloc((size_t)nnz * sizeof(edge_t<T>)));
  }
  edgelist_t(edge_t<T>* e, int _m, int _n, int _nnz) : edges(e), m(_m), n(_n), nnz(_nnz) {}
  void clear() {
    if (edges) free(edges);
    edges = nullptr;
    nnz = 0;
    m = 0;
    n = 0;
  }
};

namespace detail {
template <typename T>
bool read_text_edge(FILE* f, int* s, int* d, T* v, bool weights) {
  if (!weights) {
    *v = (T)1;
    return fscanf(f, "%d %d", s, d) == 2;
  }
  double w;
  if (fscanf(f, "%d %d %lf", s, d, &w) != 3) return false;
  *v = (T)w;
  return true;
}
}  // namespace detail

// dir is a file-name PREFIX: rank r reads <dir>r, <dir>(r + nrank), ... (edgelist.h:250-253); one rank here.
template <typename T>
void load_edgelist(const char* dir, edgelist_t<T>* edgelist, bool binaryformat = true, bool header = true,
                   bool edgeweights = true) {
  edgelist->m = edgelist->n = edgelist->nnz = 0;
  edgelist->edges = nullptr;
  size_t cap = 0;
  for (int i = 0;; i++) {
    std::stringstream name;
    name << dir << i;
    FILE* fp = fopen(name.str().c_str(), binaryformat ? "rb" : "r");
    if (!fp) {
      if (i == 0) printf("Could not open file: %s\n", name.str().c_str());
      break;
    }
    printf("Reading file: %s\n", name.str().c_str());
    int m = 0, n = 0, nnz = -1;
    if (header) {
      if (binaryformat) {
        int h[3];
        if (fread(h, sizeof(int), 3, fp) != 3) { fclose(fp); break; }
        m = h[0]; n = h[1]; nnz = h[2];
      } else {
        if (fscanf(fp, "%d %d %d", &m, &n, &nnz) != 3) { fclose(fp); break; }
      }
    }
    int s, d;
    T v;
    long long got = 0;
    while (nnz < 0 || got < nnz) {
      bool ok;
      if (binaryformat) {
        ok = fread(&s, sizeof(int), 1, fp) == 1 && fread(&d, sizeof(int), 1, fp) == 1;
        if (ok && edgeweights) ok = fread(&v, sizeof(T), 1, fp) == 1;
        if (ok && !edgeweights) v = (T)1;
      } else {
        ok = detail::read_text_edge<T>(fp, &s, &d, &v, edgeweights);
      }
      if (!ok) break;
      if ((size_t)edgelist->nnz == cap) {
        cap = cap ? cap * 2 : 1024;
        edgelist->edges = reinterpret_cast<edge_t<T>*>(realloc(edgelist->edges, cap * sizeof(edge_t<T>)));
      }
      edgelist->edges[edgelist->nnz++] = edge_t<T>(s, d, v);
      if (!header) { m = std::max(m, s); n = std::max(n, d); }
      got++;
    }
    edgelist->m = std::max(edgelist->m, m);
    edgelist->n = std::max(edgelist->n, n);
    fclose(fp);
  }
  std::cout << "Got: " << edgelist->m << " by " << edgelist->n << "  vertices" << std::endl;
  std::cout << "Got: " << edgelist->nnz << " edges" << std::endl;
}

template <typename T>
void write_edgelist(const char* dir, const edgelist_t<T>& edgelist, bool binaryformat = true, bool header = true,
                    bool edgeweights = true) {
  std::stringstream name;
  name << dir << 0;
  FILE* fp = fopen(name.str().c_str(), binaryformat ? "wb" : "w");
  if (!fp) { printf("Could not open file: %s\n", name.str().c_str()); return; }
  if (header) {
    if (binaryformat) { int h[3] = {edgelist.m, edgelist.n, edgelist.nnz}; fwrite(h, sizeof(int), 3, fp); }
    else fprintf(fp, "%d %d %d\n", edgelist.m, edgelist.n, edgelist.nnz);
  }
  for (int i = 0; i < edgelist.nnz; i++) {
    const edge_t<T>& e = edgelist.edges[i];
    if (binaryformat) {
      fwrite(&e.src, sizeof(int), 1, fp);
      fwrite(&e.dst, sizeof(int), 1, fp);
      if (edgeweights) fwrite(&e.val, sizeof(T), 1, fp);
    } else if (edgeweights) {
      fprintf(fp, "%d %d %.9g\n", e.src, e.dst, (double)e.val);
    } else {
      fprintf(fp, "%d %d\n", e.src, e.dst);
    }
  }
  fclose(fp);
}

// edgelist_transformation.h:431-443
template <typename T>
edgelist_t<T> filter_edges(edgelist_t<T>* edgelist, bool (*filter_function)(edge_t<T>, void*), void* param = NULL) {
  edgelist_t<T> out(edgelist->m, edgelist->n, edgelist->nnz);
  int k = 0;
  for (int i = 0; i < edgelist->nnz; i++)
    if (filter_function(edgelist->edges[i], param)) out.edges[k++] = edgelist->edges[i];
  out.nnz = k;
  return out;
}

}  // namespace GraphMat
#endif
