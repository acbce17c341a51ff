/* Host-side input decoding for the reference's TFRecord data format (README.md:49-64; parser at
 * source_segmenter.py:331-355 / adversarial.py:607-631) without TensorFlow: record framing (length, masked CRC32C,
 * payload, masked CRC32C), the tf.train.Example protobuf wire format, tf.decode_raw / reshape / slice -- in plain C so
 * that the Python reader threads (tfrecord.py) run it with the GIL released.  One example = two raw float32 volumes
 * of 256x256x3 (1.57 MB); the 8-GPU adversarial step consumes ~5 GB/s of them.
 *
 * Built by _build.py with gcc into libpnp_io.so (no CUDA dependency); C-ABI in include/pnp_io.h. */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pnp_io.h"

/* ---- CRC32C (Castagnoli): SSE4.2 crc32 instruction when the CPU has it, slicing-by-8 tables otherwise ---------------- */
static uint32_t g_tab[8][256];
static int g_tab_ready = 0;

static void init_tables(void) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xFF];
  g_tab_ready = 1;
}

static uint32_t crc_sw(uint32_t c, const uint8_t* p, size_t n) {
  if (!g_tab_ready) init_tables();
  while (n && ((uintptr_t)p & 7)) { c = g_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= c;
    c = g_tab[7][v & 0xFF] ^ g_tab[6][(v >> 8) & 0xFF] ^ g_tab[5][(v >> 16) & 0xFF] ^ g_tab[4][(v >> 24) & 0xFF] ^
        g_tab[3][(v >> 32) & 0xFF] ^ g_tab[2][(v >> 40) & 0xFF] ^ g_tab[1][(v >> 48) & 0xFF] ^ g_tab[0][(v >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) static uint32_t crc_hw(uint32_t c, const uint8_t* p, size_t n) {
  uint64_t c64 = c;
  while (n && ((uintptr_t)p & 7)) { c64 = __builtin_ia32_crc32qi((uint32_t)c64, *p++); --n; }
  /* three independent streams hide the 3-cycle latency of crc32q; combining them needs a carry-less shift, so keep it simple:
     one stream, 8 bytes per instruction (measured ~8 GB/s per core, far above one reader thread's share) */
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    c64 = __builtin_ia32_crc32di(c64, v);
    p += 8;
    n -= 8;
  }
  while (n--) c64 = __builtin_ia32_crc32qi((uint32_t)c64, *p++);
  return (uint32_t)c64;
}
static int have_hw(void) {
  static int known = -1;
  if (known < 0) known = __builtin_cpu_supports("sse4.2") ? 1 : 0;
  return known;
}
#else
static uint32_t crc_hw(uint32_t c, const uint8_t* p, size_t n) { return crc_sw(c, p, n); }
static int have_hw(void) { return 0; }
#endif

uint32_t pnp_crc32c(const uint8_t* data, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  c = have_hw() ? crc_hw(c, data, n) : crc_sw(c, data, n);
  return c ^ 0xFFFFFFFFu;
}

uint32_t pnp_crc32c_sw(const uint8_t* data, size_t n) { return crc_sw(0xFFFFFFFFu, data, n) ^ 0xFFFFFFFFu; }

uint32_t pnp_masked_crc32c(const uint8_t* data, size_t n) {
  uint32_t c = pnp_crc32c(data, n);
  return (uint32_t)(((c >> 15) | (c << 17)) + 0xA282EAD8u);
}

int pnp_crc32c_is_hardware(void) { return have_hw(); }

/* ---- protobuf wire format ------------------------------------------------------------------------------------------------ */
static int varint(const uint8_t* b, size_t n, size_t* pos, uint64_t* out) {
  uint64_t v = 0;
  int shift = 0;
  while (*pos < n && shift < 64) {
    uint8_t x = b[(*pos)++];
    v |= (uint64_t)(x & 0x7F) << shift;
    if (!(x & 0x80)) { *out = v; return 0; }
    shift += 7;
  }
  return -1;
}

/* next field of a message: number, wire type, and for length-delimited fields the (ptr, len) of the value */
static int next_field(const uint8_t* b, size_t n, size_t* pos, uint32_t* fn, uint32_t* wt, const uint8_t** val, size_t* len,
                      uint64_t* ival) {
  uint64_t key;
  if (varint(b, n, pos, &key)) return -1;
  *fn = (uint32_t)(key >> 3);
  *wt = (uint32_t)(key & 7);
  *val = NULL; *len = 0; *ival = 0;
  if (*wt == 0) return varint(b, n, pos, ival);
  if (*wt == 2) {
    uint64_t l;
    if (varint(b, n, pos, &l) || l > n - *pos) return -1;
    *val = b + *pos; *len = (size_t)l; *pos += (size_t)l;
    return 0;
  }
  if (*wt == 1) { if (n - *pos < 8) return -1; *val = b + *pos; *len = 8; *pos += 8; return 0; }
  if (*wt == 5) { if (n - *pos < 4) return -1; *val = b + *pos; *len = 4; *pos += 4; return 0; }
  return -1;
}

/* Example{1: Features{1: map entry{1: key, 2: Feature{1: BytesList{1: bytes}}}}} -> the single bytes value of feature `name` */
static int find_bytes_feature(const uint8_t* ex, size_t n, const char* name, const uint8_t** out, size_t* out_len) {
  size_t p0 = 0, nl = strlen(name);
  uint32_t fn, wt; const uint8_t* v; size_t l; uint64_t iv;
  while (p0 < n) {
    if (next_field(ex, n, &p0, &fn, &wt, &v, &l, &iv)) return PNP_IO_ERR_PROTO;
    if (fn != 1 || wt != 2) continue;
    const uint8_t* feats = v; size_t fl = l, p1 = 0;
    while (p1 < fl) {
      if (next_field(feats, fl, &p1, &fn, &wt, &v, &l, &iv)) return PNP_IO_ERR_PROTO;
      if (fn != 1 || wt != 2) continue;
      const uint8_t* entry = v; size_t el = l, p2 = 0;
      const uint8_t* key = NULL; size_t kl = 0; const uint8_t* feat = NULL; size_t ftl = 0;
      while (p2 < el) {
        if (next_field(entry, el, &p2, &fn, &wt, &v, &l, &iv)) return PNP_IO_ERR_PROTO;
        if (fn == 1 && wt == 2) { key = v; kl = l; }
        else if (fn == 2 && wt == 2) { feat = v; ftl = l; }
      }
      if (!key || kl != nl || memcmp(key, name, nl) != 0 || !feat) continue;
      size_t p3 = 0;
      while (p3 < ftl) {
        if (next_field(feat, ftl, &p3, &fn, &wt, &v, &l, &iv)) return PNP_IO_ERR_PROTO;
        if (fn != 1 || wt != 2) continue;                 /* 1 = BytesList */
        const uint8_t* lst = v; size_t ll = l, p4 = 0;
        while (p4 < ll) {
          if (next_field(lst, ll, &p4, &fn, &wt, &v, &l, &iv)) return PNP_IO_ERR_PROTO;
          if (fn == 1 && wt == 2) { *out = v; *out_len = l; return 0; }
        }
      }
      return PNP_IO_ERR_SCHEMA;
    }
  }
  return PNP_IO_ERR_SCHEMA;
}

/* ---- record framing --------------------------------------------------------------------------------------------------------- */
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

static int locate_record(const uint8_t* buf, size_t n, int index, int check_crc, const uint8_t** payload, size_t* plen) {
  size_t pos = 0;
  int k = 0;
  while (pos + 12 <= n) {
    uint64_t len = rd64(buf + pos);
    if (check_crc && pnp_masked_crc32c(buf + pos, 8) != rd32(buf + pos + 8)) return PNP_IO_ERR_CRC;
    if (len > n - pos - 16) return PNP_IO_ERR_TRUNCATED;
    const uint8_t* data = buf + pos + 12;
    if (k == index) {
      if (check_crc && pnp_masked_crc32c(data, (size_t)len) != rd32(data + len)) return PNP_IO_ERR_CRC;
      *payload = data; *plen = (size_t)len;
      return 0;
    }
    pos += 12 + (size_t)len + 4;
    ++k;
  }
  return index < 0 ? k : PNP_IO_ERR_INDEX;     /* index < 0: count the records */
}

int pnp_tfrecord_count(const uint8_t* buf, size_t n) {
  const uint8_t* p; size_t l;
  return locate_record(buf, n, -1, 0, &p, &l);
}

int pnp_tfrecord_decode(const uint8_t* buf, size_t n, int record_index, int check_crc, float* image, long long* label, int H, int W,
                        int C, int label_channel) {
  if (!buf || !image || !label || H <= 0 || W <= 0 || C <= 0 || label_channel < 0 || label_channel >= C || record_index < 0)
    return PNP_IO_ERR_ARG;
  const uint8_t* ex; size_t exl;
  int rc = locate_record(buf, n, record_index, check_crc, &ex, &exl);
  if (rc) return rc;
  const size_t vol_bytes = (size_t)H * W * C * 4;
  const uint8_t* dv; size_t dl; const uint8_t* lv; size_t ll;
  if ((rc = find_bytes_feature(ex, exl, "data_vol", &dv, &dl))) return rc;
  if ((rc = find_bytes_feature(ex, exl, "label_vol", &lv, &ll))) return rc;
  if (dl != vol_bytes || ll != vol_bytes) return PNP_IO_ERR_SCHEMA;
  memcpy(image, dv, vol_bytes);                       /* tf.decode_raw(float32) + reshape + slice [0,0,0]:[H,W,C] */
  const size_t px = (size_t)H * W;
  for (size_t i = 0; i < px; ++i) {                   /* tf.slice(label_vol, [0,0,1], [H,W,1]) -> integer class map */
    float f;
    memcpy(&f, lv + (i * C + label_channel) * 4, 4);
    label[i] = (long long)f;
  }
  return 0;
}

int pnp_tfrecord_load_file(const char* path, int record_index, int check_crc, float* image, long long* label, int H, int W, int C,
                           int label_channel) {
  FILE* f = fopen(path, "rb");
  if (!f) return PNP_IO_ERR_OPEN;
  setvbuf(f, NULL, _IONBF, 0);            /* one large fread straight into the staging buffer */
  if (fseek(f, 0, SEEK_END)) { fclose(f); return PNP_IO_ERR_OPEN; }
  long sz = ftell(f);
  if (sz < 0 || fseek(f, 0, SEEK_SET)) { fclose(f); return PNP_IO_ERR_OPEN; }
  /* one staging buffer per reader thread, grown on demand and kept: a fresh 1.5 MB malloc per example is an mmap + page
     faults + munmap, which cost more than the decode itself */
  static __thread uint8_t* tl_buf = NULL;
  static __thread size_t tl_cap = 0;
  if ((size_t)sz > tl_cap) {
    free(tl_buf);
    tl_cap = ((size_t)sz + (1u << 20)) & ~(((size_t)1 << 20) - 1);
    tl_buf = (uint8_t*)malloc(tl_cap);
    if (!tl_buf) { tl_cap = 0; fclose(f); return PNP_IO_ERR_OPEN; }
  }
  size_t got = fread(tl_buf, 1, (size_t)sz, f);
  fclose(f);
  return (got == (size_t)sz) ? pnp_tfrecord_decode(tl_buf, got, record_index, check_crc, image, label, H, W, C, label_channel)
                             : PNP_IO_ERR_TRUNCATED;
}

const char* pnp_io_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case PNP_IO_ERR_ARG: return "pnp_io: bad argument";
    case PNP_IO_ERR_OPEN: return "pnp_io: cannot open / read file";
    case PNP_IO_ERR_TRUNCATED: return "pnp_io: truncated record";
    case PNP_IO_ERR_CRC: return "pnp_io: corrupt record (masked CRC32C mismatch)";
    case PNP_IO_ERR_PROTO: return "pnp_io: malformed tf.train.Example";
    case PNP_IO_ERR_SCHEMA: return "pnp_io: example does not follow the data_vol / label_vol schema (README.md:49-64)";
    case PNP_IO_ERR_INDEX: return "pnp_io: record index out of range";
    default: return "pnp_io: unknown error";
  }
}
