/* C-ABI of libpnp_io.so: host-side decoding of the reference's TFRecord input format -- what TensorFlow's
 * TFRecordReader / parse_single_example / decode_raw do for source_segmenter.py:331-355 and adversarial.py:607-631
 * (schema: README.md:49-64).  Plain C, no CUDA; called from the reader threads of tfrecord.py with the GIL released.
 * Every function returns 0 on success or a negative PNP_IO_ERR_* code; nothing throws, nothing allocates that the
 * caller must free. */
#ifndef PNP_IO_H
#define PNP_IO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PNP_IO_ERR_ARG (-1)
#define PNP_IO_ERR_OPEN (-2)
#define PNP_IO_ERR_TRUNCATED (-3)
#define PNP_IO_ERR_CRC (-4)
#define PNP_IO_ERR_PROTO (-5)
#define PNP_IO_ERR_SCHEMA (-6)
#define PNP_IO_ERR_INDEX (-7)

/* CRC32C (Castagnoli) of a byte range -- SSE4.2 crc32 instruction when present, slicing-by-8 otherwise (pnp_crc32c_sw:
 * always the table path, for the self-check) -- and TFRecord's masked form ((crc >> 15 | crc << 17) + 0xa282ead8) */
uint32_t pnp_crc32c(const uint8_t* data, size_t n);
uint32_t pnp_crc32c_sw(const uint8_t* data, size_t n);
uint32_t pnp_masked_crc32c(const uint8_t* data, size_t n);
int pnp_crc32c_is_hardware(void);

/* number of records framed in a TFRecord byte image (>= 0), or a negative error */
int pnp_tfrecord_count(const uint8_t* buf, size_t n);
/* Decode record `record_index` of a TFRecord byte image: image[H*W*C] = data_vol (tf.decode_raw float32 + reshape),
 * label[H*W] = (int64) label_vol[:, :, label_channel] (tf.slice(label_vol, [0,0,1], [H,W,1]): the MIDDLE slice for the
 * reference's label_channel = 1).  check_crc != 0 verifies both masked CRCs of the record. */
int pnp_tfrecord_decode(const uint8_t* buf, size_t n, int record_index, int check_crc, float* image, long long* label, int H, int W,
                        int C, int label_channel);
/* the same for a file on disk (the reference's lists/.._list name single-example files) */
int pnp_tfrecord_load_file(const char* path, int record_index, int check_crc, float* image, long long* label, int H, int W, int C,
                           int label_channel);
const char* pnp_io_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif
