// edgelist.h -- host-side mirror of the reference's edge-list types and loaders
// and edge-list transformations (narayanan2004/GraphMat include/GMDP/utils/edgelist.h:38-78,242-454 and
// edgelist_transformation.h:37-443): same names and argument meaning, single process.
// File format (binary): int m, n, nnz header, then (int src, int dst, T val) records,
// ids 1-based.  The header's nnz is authoritative (the reference reads to EOF into a buffer
// sized from the header and overruns by one record on the shipped data files, SURVEY.md
// hazard 6); records beyond it are ignored here.
#ifndef GRAPHMAT_B200_EDGELIST_H
#define GRAPHMAT_B200_EDGELIST_H
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
// Text fields are parsed and printed in the weight's own type, with the reference's conversions
// (edgelist.h:89-131 readLine, 176-206 writeLine): a float is read with %f, never through a double.
template <typename T> struct text_fmt;
template <> struct text_fmt<int>          { static const char* in() { return "%d"; }  static const char* out() { return "%d"; } };
template <> struct text_fmt<unsigned int> { static const char* in() { return "%u"; }  static const char* out() { return "%u"; } };
template <> struct text_fmt<float>        { static const char* in() { return "%f"; }  static const char* out() { return "%.8f"; } };
template <> struct text_fmt<double>       { static const char* in() { return "%lf"; } static const char* out() { return "%.15lf"; } };

template <typename T>
bool read_text_edge(FILE* f, int* s, int* d, T* v, bool weights) {
  if (fscanf(f, "%d %d", s, d) != 2) return false;
  if (!weights) { *v = (T)1; return true; }
  return fscanf(f, text_fmt<T>::in(), v) == 1;
}
// grows the array to hold at least `need` edges (doubling; exact when the count is known up front)
template <typename T>
void reserve(edgelist_t<T>* el, size_t* cap, size_t need, bool exact = false) {
  if (need <= *cap) return;
  size_t c = exact ? need : std::max<size_t>(need, *cap ? *cap * 2 : 1024);
  el->edges = reinterpret_cast<edge_t<T>*>(realloc(el->edges, c * sizeof(edge_t<T>)));
  *cap = c;
}

// Binary records are (int src, int dst[, T val]) back to back.  With weights that is the memory layout of edge_t<T>
// for 4- and 8-byte T, so the file is read straight into the array in one fread per file (the reference reads field
// by field, edgelist.h:89-106: three library calls per edge, ~40 s for a billion edges); without weights the id pairs
// are read in blocks and widened.  nnz < 0: no header, the record count comes from the file length and m, n from
// the ids.  A header that promises more records than the file holds gets what is there.
template <typename T>
bool read_binary_records(FILE* fp, edgelist_t<T>* el, size_t* cap, long long nnz, bool weights, bool header, int* m, int* n) {
  const size_t rec = 2 * sizeof(int) + (weights ? sizeof(T) : 0);
  const long long here = ftello(fp);
  if (here < 0 || fseeko(fp, 0, SEEK_END)) return false;
  const long long in_file = (ftello(fp) - here) / (long long)rec;
  if (fseeko(fp, here, SEEK_SET)) return false;
  const long long want = nnz < 0 ? in_file : std::min(nnz, in_file);
  if (want <= 0) return true;
  if ((long long)el->nnz + want > 0x7fffffffLL) {
    printf("graphmat_b200: more than 2^31-1 edges do not fit edgelist_t (int nnz, as in the reference)\n");
    return false;
  }
  reserve(el, cap, (size_t)el->nnz + (size_t)want, true);
  edge_t<T>* out = el->edges + el->nnz;
  long long got = 0;
  if (weights && sizeof(edge_t<T>) == rec) {
    got = (long long)fread(out, rec, (size_t)want, fp);
  } else {
    const size_t block = 1 << 20;
    unsigned char* buf = reinterpret_cast<unsigned char*>(malloc(block * rec));
    while (got < want) {
      const size_t k = fread(buf, rec, (size_t)std::min<long long>((long long)block, want - got), fp);
      if (k == 0) break;
      for (size_t i = 0; i < k; i++) {
        const unsigned char* r = buf + i * rec;
        edge_t<T>& e = out[got + (long long)i];
        memcpy(&e.src, r, sizeof(int));
        memcpy(&e.dst, r + sizeof(int), sizeof(int));
        if (weights) memcpy(&e.val, r + 2 * sizeof(int), sizeof(T));
        else e.val = (T)1;
      }
      got += (long long)k;
    }
    free(buf);
  }
  if (!header)
    for (long long i = 0; i < got; i++) { *m = std::max(*m, out[i].src); *n = std::max(*n, out[i].dst); }
  el->nnz += (int)got;
  return true;
}

// the writing side of the same layouts
template <typename T>
void write_binary_records(FILE* fp, const edgelist_t<T>& el, bool weights) {
  const size_t rec = 2 * sizeof(int) + (weights ? sizeof(T) : 0);
  if (weights && sizeof(edge_t<T>) == rec) {
    fwrite(el.edges, rec, (size_t)el.nnz, fp);
    return;
  }
  const size_t block = 1 << 20;
  unsigned char* buf = reinterpret_cast<unsigned char*>(malloc(block * rec));
  for (size_t done = 0; done < (size_t)el.nnz;) {
    const size_t k = std::min(block, (size_t)el.nnz - done);
    for (size_t i = 0; i < k; i++) {
      const edge_t<T>& e = el.edges[done + i];
      unsigned char* r = buf + i * rec;
      memcpy(r, &e.src, sizeof(int));
      memcpy(r + sizeof(int), &e.dst, sizeof(int));
      if (weights) memcpy(r + 2 * sizeof(int), &e.val, sizeof(T));
    }
    fwrite(buf, rec, k, fp);
    done += k;
  }
  free(buf);
}

template <typename T>
void write_text_edge(FILE* f, int s, int d, const T& v, bool weights) {
  fprintf(f, "%d %d", s, d);
  if (weights) { fputc(' ', f); fprintf(f, text_fmt<T>::out(), v); }
  fputc('\n', f);
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
    if (binaryformat) {
      if (!detail::read_binary_records<T>(fp, edgelist, &cap, nnz, edgeweights, header, &m, &n)) { fclose(fp); break; }
    } else {
      int s, d;
      T v;
      long long got = 0;
      while (nnz < 0 || got < nnz) {
        if (!detail::read_text_edge<T>(fp, &s, &d, &v, edgeweights)) break;
        detail::reserve(edgelist, &cap, (size_t)edgelist->nnz + 1);
        edgelist->edges[edgelist->nnz++] = edge_t<T>(s, d, v);
        if (!header) { m = std::max(m, s); n = std::max(n, d); }
        got++;
      }
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
  printf("Writing file: %s\n", name.str().c_str());
  if (header) {
    if (binaryformat) { int h[3] = {edgelist.m, edgelist.n, edgelist.nnz}; fwrite(h, sizeof(int), 3, fp); }
    else fprintf(fp, "%d %d %d\n", edgelist.m, edgelist.n, edgelist.nnz);
  }
  if (binaryformat) {
    detail::write_binary_records<T>(fp, edgelist, edgeweights);
  } else {
    for (int i = 0; i < edgelist.nnz; i++) {
      const edge_t<T>& e = edgelist.edges[i];
      detail::write_text_edge<T>(fp, e.src, e.dst, e.val, edgeweights);
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

// ---- edge-list transformations (edgelist_transformation.h:37-429, edgelist.h:336-365) -------------------
// What graph_converter applies between reading and writing.  One process: the reference's shuffle_edges
// (redistribution by src % nranks) is the identity here and is not restated.

namespace detail {
template <typename T>
void adopt(edgelist_t<T>* into, edge_t<T>* edges, size_t nnz) {  // keeps m and n
  if (into->edges) free(into->edges);
  into->edges = edges;
  into->nnz = (int)nnz;
}
template <typename T>
edge_t<T>* alloc_edges(size_t n) {
  return reinterpret_cast<edge_t<T>*>(malloc((n ? n : 1) * sizeof(edge_t<T>)));
}
}  // namespace detail

// Drops (v, v) edges; order kept (edgelist_transformation.h:38-53).
template <typename T>
void remove_selfedges(edgelist_t<T>* edgelist) {
  edge_t<T>* e = edgelist->edges;
  edge_t<T>* end = std::remove_if(e, e + edgelist->nnz, [](const edge_t<T>& x) { return x.src == x.dst; });
  edgelist->nnz = (int)(end - e);
}

// One edge per (src, dst), output sorted by (src, dst) (edgelist_transformation.h:70-93,238-250).  The
// reference sorts with an unstable parallel sort, so WHICH duplicate's weight survives is unspecified
// there; here it is the first one in input order.
template <typename T>
void remove_duplicate_edges(edgelist_t<T>* edgelist) {
  if (edgelist->nnz <= 0) return;
  edge_t<T>* e = edgelist->edges;
  std::stable_sort(e, e + edgelist->nnz, [](const edge_t<T>& a, const edge_t<T>& b) {
    return a.src != b.src ? a.src < b.src : a.dst < b.dst;
  });
  edge_t<T>* end = std::unique(e, e + edgelist->nnz,
                               [](const edge_t<T>& a, const edge_t<T>& b) { return a.src == b.src && a.dst == b.dst; });
  edgelist->nnz = (int)(end - e);
}

// (u, v, w) -> (u, v, w), (v, u, w), interleaved (edgelist_transformation.h:397-410).
template <typename T>
void create_bidirectional_edges(edgelist_t<T>* edgelist) {
  size_t nnz = (size_t)edgelist->nnz;
  edge_t<T>* out = detail::alloc_edges<T>(2 * nnz);
  for (size_t i = 0; i < nnz; i++) {
    const edge_t<T>& e = edgelist->edges[i];
    out[2 * i] = e;
    out[2 * i + 1] = edge_t<T>(e.dst, e.src, e.val);
  }
  detail::adopt(edgelist, out, 2 * nnz);
}

// Every edge points from the smaller id to the larger (edgelist_transformation.h:413-419).
template <typename T>
void convert_to_dag(edgelist_t<T>* edgelist) {
  for (int i = 0; i < edgelist->nnz; i++) {
    edge_t<T>& e = edgelist->edges[i];
    if (e.src > e.dst) std::swap(e.src, e.dst);
  }
}

// Weights drawn with the C library's rand() in edge order, clamped to [1, range]
// (edgelist_transformation.h:422-429): the same sequence as the reference on the same libc.
template <typename T>
void random_edge_weights(edgelist_t<T>* edgelist, int random_range) {
  for (int i = 0; i < edgelist->nnz; i++) {
    double t = (double)rand() / (double)RAND_MAX * (double)random_range;
    t = std::min(std::max(t, 1.0), (double)random_range);
    edgelist->edges[i].val = (T)t;
  }
}

// Coin flip per edge (edgelist_transformation.h:388-394).
template <typename T>
void randomize_edge_direction(edgelist_t<T>* edgelist) {
  for (int i = 0; i < edgelist->nnz; i++)
    if ((double)rand() / (double)RAND_MAX < 0.5) std::swap(edgelist->edges[i].src, edgelist->edges[i].dst);
}

// Relabels the vertices of a square edge list with the permutation the reference draws
// (edgelist.h:336-365): srand(5), one rand() % m per vertex drawn up front, then position i is
// exchanged with its drawn partner for i = 0 .. m-1.
template <typename T>
void randomize_edgelist_square(edgelist_t<T>* edgelist) {
  const int m = edgelist->m;
  unsigned* label = new unsigned[m > 0 ? m : 1];
  unsigned* partner = new unsigned[m > 0 ? m : 1];
  srand(5);
  for (int i = 0; i < m; i++) {
    label[i] = (unsigned)i;
    partner[i] = (unsigned)(rand() % m);
  }
  for (int i = 0; i < m; i++) std::swap(label[i], label[partner[i]]);
  for (int i = 0; i < edgelist->nnz; i++) {
    edge_t<T>& e = edgelist->edges[i];
    e.src = (int)label[e.src - 1] + 1;
    e.dst = (int)label[e.dst - 1] + 1;
  }
  delete[] partner;
  delete[] label;
}

// Largest source and destination id (edgelist.h:423-436).
template <typename T>
void get_dimensions(const edge_t<T>* edges, int nnz, int& max_m, int& max_n) {
  max_m = max_n = 0;
  for (int i = 0; i < nnz; i++) {
    max_m = std::max(max_m, edges[i].src);
    max_n = std::max(max_n, edges[i].dst);
  }
}

// Keeps rows [start_row, end_row) (0-based) and renumbers them from 1 (edgelist.h:403-420).
template <typename T>
void filter_edges_by_row(edgelist_t<T>* edges, int start_row, int end_row) {
  int kept = 0;
  for (int i = 0; i < edges->nnz; i++) {
    edge_t<T> e = edges->edges[i];
    if (e.src - 1 < start_row || e.src - 1 >= end_row) continue;
    e.src -= start_row;
    edges->edges[kept++] = e;
  }
  edges->nnz = kept;
  edges->m = end_row - start_row;
}

// Renumbers the destination ids over the columns that occur; *remaining_indices (new[]-allocated) lists
// the surviving original ids, 1-based (edgelist.h:367-401).
template <typename T>
void remove_empty_columns(edgelist_t<T>* edges, int** remaining_indices) {
  const int n = edges->n;
  int* rank = new int[n + 1]();
  for (int i = 0; i < edges->nnz; i++) rank[edges->edges[i].dst] = 1;  // rank[c] for 1-based column c
  int kept = 0;
  for (int c = 1; c <= n; c++) kept += rank[c];
  *remaining_indices = new int[kept > 0 ? kept : 1];
  int next = 0;
  for (int c = 1; c <= n; c++) {
    if (!rank[c]) continue;
    (*remaining_indices)[next] = c;
    rank[c] = ++next;
  }
  for (int i = 0; i < edges->nnz; i++) edges->edges[i].dst = rank[edges->edges[i].dst];
  edges->n = kept;
  delete[] rank;
}

// edgelist.h:439-454
template <typename T>
void ReadEdges(edgelist_t<T>* edgelist, const char* fname_in, bool binaryformat = true, bool header = true,
               bool edgeweights = true, bool randomize = false) {
  load_edgelist(fname_in, edgelist, binaryformat, header, edgeweights);
  if (randomize) randomize_edgelist_square<T>(edgelist);
}
template <typename T>
void WriteEdges(const edgelist_t<T>& edgelist, const char* fname_in, bool binaryformat = true, bool header = true,
                bool edgeweights = true) {
  write_edgelist(fname_in, edgelist, binaryformat, header, edgeweights);
}

}  // namespace GraphMat
#endif
