/* TEST INFRASTRUCTURE ONLY (oracle/ref_harness: the recipe that pins the oracle against the UNMODIFIED reference on a box
 * that has gfortran).  A stand-in for the nine HDF4 SD entry points the reference driver calls (equiSources.f90:56,
 * :316-423 grid input, :1087-1154 restart, :4843-4905 cell-array output), so that `program pointTransfer` links without
 * libmfhdf / libdf.  Datasets live in a flat container file:
 *     "RTBSD001" | int32 nsds | nsds x { char name[64] | int32 type | int32 rank | int32 dims[4] | int64 nbytes | data }
 * (dims in Fortran order, data exactly as the Fortran array lies in memory).  radiativetransfer_b200/formats.py reads
 * and writes the same container (write_sd_container / read_sd_container).
 * Calling convention: gfortran -- lower case + trailing underscore, arguments by reference, hidden string lengths by
 * value at the end of the argument list.  Only whole-array transfers (start = 0, stride = 1), which is all the driver does. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_FILES 8
#define MAX_SDS 256

typedef struct {
  char name[64];
  int32_t type, rank, dims[4];
  int64_t nbytes;
  void* data;
} Sds;

typedef struct {
  int used, writing;
  char path[1024];
  int nsds;
  Sds sds[MAX_SDS];
} SdFile;

static SdFile g_files[MAX_FILES];

static int elem_size(int type) {
  switch (type) {
    case 3: case 4: case 20: case 21: return 1;
    case 22: case 23: return 2;
    case 5: case 24: case 25: return 4;
    case 6: return 8;
    default: return 0;
  }
}

static void copy_fstring(char* dst, size_t cap, const char* src, long len) {
  while (len > 0 && src[len - 1] == ' ') len--;
  if ((size_t)len >= cap) len = (long)cap - 1;
  memcpy(dst, src, (size_t)len);
  dst[len] = 0;
}

/* sd_id = file slot + 1; sds_id = (file slot + 1) * 1000 + dataset index */
int sfstart_(const char* name, const int* access, long name_len) {
  int slot = -1;
  for (int i = 0; i < MAX_FILES; i++)
    if (!g_files[i].used) { slot = i; break; }
  if (slot < 0) return -1;
  SdFile* f = &g_files[slot];
  memset(f, 0, sizeof(*f));
  copy_fstring(f->path, sizeof(f->path), name, name_len);
  f->used = 1;
  f->writing = (*access == 4);   /* dfacc_create */
  if (!f->writing) {
    FILE* fp = fopen(f->path, "rb");
    char magic[8];
    int32_t n = 0;
    if (!fp || fread(magic, 1, 8, fp) != 8 || memcmp(magic, "RTBSD001", 8) != 0 || fread(&n, 4, 1, fp) != 1 || n < 0 || n > MAX_SDS) {
      fprintf(stderr, "hdf4_stub: cannot read container %s\n", f->path);
      if (fp) fclose(fp);
      f->used = 0;
      return -1;
    }
    f->nsds = n;
    for (int i = 0; i < n; i++) {
      Sds* s = &f->sds[i];
      if (fread(s->name, 1, 64, fp) != 64 || fread(&s->type, 4, 1, fp) != 1 || fread(&s->rank, 4, 1, fp) != 1 ||
          fread(s->dims, 4, 4, fp) != 4 || fread(&s->nbytes, 8, 1, fp) != 1) { fclose(fp); f->used = 0; return -1; }
      s->data = malloc((size_t)(s->nbytes > 0 ? s->nbytes : 1));
      if (!s->data || fread(s->data, 1, (size_t)s->nbytes, fp) != (size_t)s->nbytes) { fclose(fp); f->used = 0; return -1; }
    }
    fclose(fp);
  }
  return slot + 1;
}

int sffinfo_(const int* sd_id, int* n_datasets, int* n_file_attrs) {
  if (*sd_id < 1 || *sd_id > MAX_FILES || !g_files[*sd_id - 1].used) return -1;
  *n_datasets = g_files[*sd_id - 1].nsds;
  *n_file_attrs = 0;
  return 0;
}

int sfselect_(const int* sd_id, const int* index) {
  if (*sd_id < 1 || *sd_id > MAX_FILES || !g_files[*sd_id - 1].used) return -1;
  if (*index < 0 || *index >= g_files[*sd_id - 1].nsds) return -1;
  return *sd_id * 1000 + *index;
}

static Sds* find_sds(int sds_id) {
  const int slot = sds_id / 1000 - 1, idx = sds_id % 1000;
  if (slot < 0 || slot >= MAX_FILES || !g_files[slot].used || idx >= g_files[slot].nsds) return NULL;
  return &g_files[slot].sds[idx];
}

int sfginfo_(const int* sds_id, char* name, int* rank, int* dims, int* type, int* nattrs, long name_len) {
  Sds* s = find_sds(*sds_id);
  if (!s) return -1;
  memset(name, ' ', (size_t)name_len);
  size_t n = strlen(s->name);
  if ((long)n > name_len) n = (size_t)name_len;
  memcpy(name, s->name, n);
  *rank = s->rank;
  for (int i = 0; i < s->rank; i++) dims[i] = s->dims[i];
  *type = s->type;
  *nattrs = 0;
  return 0;
}

static int64_t edges_bytes(const Sds* s, const int* start, const int* stride, const int* edges) {
  int64_t n = 1;
  for (int i = 0; i < s->rank; i++) {
    if (start[i] != 0 || stride[i] != 1) return -1;   /* whole arrays only */
    n *= edges[i];
  }
  return n * elem_size(s->type);
}

int sfrdata_(const int* sds_id, const int* start, const int* stride, const int* edges, void* data) {
  Sds* s = find_sds(*sds_id);
  if (!s) return -1;
  const int64_t nb = edges_bytes(s, start, stride, edges);
  if (nb < 0 || nb > s->nbytes) { fprintf(stderr, "hdf4_stub: partial / oversized read of %s\n", s->name); return -1; }
  memcpy(data, s->data, (size_t)nb);
  return 0;
}

int sfendacc_(const int* sds_id) { (void)sds_id; return 0; }

int sfcreate_(const int* sd_id, const char* name, const int* type, const int* rank, const int* dims, long name_len) {
  if (*sd_id < 1 || *sd_id > MAX_FILES || !g_files[*sd_id - 1].used || !g_files[*sd_id - 1].writing) return -1;
  SdFile* f = &g_files[*sd_id - 1];
  if (f->nsds >= MAX_SDS || *rank < 1 || *rank > 4 || !elem_size(*type)) return -1;
  Sds* s = &f->sds[f->nsds];
  memset(s, 0, sizeof(*s));
  copy_fstring(s->name, sizeof(s->name), name, name_len);
  s->type = *type;
  s->rank = *rank;
  for (int i = 0; i < *rank; i++) s->dims[i] = dims[i];
  return *sd_id * 1000 + f->nsds++;
}

int sfwdata_(const int* sds_id, const int* start, const int* stride, const int* edges, const void* data) {
  Sds* s = find_sds(*sds_id);
  if (!s) return -1;
  const int64_t nb = edges_bytes(s, start, stride, edges);
  if (nb < 0) return -1;
  free(s->data);
  s->data = malloc((size_t)(nb > 0 ? nb : 1));
  if (!s->data) return -1;
  memcpy(s->data, data, (size_t)nb);
  s->nbytes = nb;
  return 0;
}

int sfend_(const int* sd_id) {
  if (*sd_id < 1 || *sd_id > MAX_FILES || !g_files[*sd_id - 1].used) return -1;
  SdFile* f = &g_files[*sd_id - 1];
  int rc = 0;
  if (f->writing) {
    FILE* fp = fopen(f->path, "wb");
    if (!fp) rc = -1;
    else {
      const int32_t n = f->nsds;
      fwrite("RTBSD001", 1, 8, fp);
      fwrite(&n, 4, 1, fp);
      for (int i = 0; i < n; i++) {
        Sds* s = &f->sds[i];
        fwrite(s->name, 1, 64, fp); fwrite(&s->type, 4, 1, fp); fwrite(&s->rank, 4, 1, fp);
        fwrite(s->dims, 4, 4, fp); fwrite(&s->nbytes, 8, 1, fp); fwrite(s->data, 1, (size_t)s->nbytes, fp);
      }
      fclose(fp);
    }
  }
  for (int i = 0; i < f->nsds; i++) free(f->sds[i].data);
  f->used = 0;
  return rc;
}
