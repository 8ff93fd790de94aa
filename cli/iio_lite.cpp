// Library-free stand-in for the three `iio` entry points the reference's TV-L1 CLI uses
// (src/iio.h:28,83,245), so that the UNMODIFIED src/tvl1flow_main.cpp links and runs without
// libpng / libjpeg / libtiff (the reference's own src/iio.cpp hard-enables those, src/iio.h:267-269,
// and does not compile in this image).
//
//   reads : PGM (P5 binary, P2 ascii; 8 or 16 bit), PFM (Pf, grey, either endianness)
//   writes: Middlebury .flo for 2-channel data -- "PIEH", int32 w, int32 h, interleaved float32 (u,v)
//           exactly as src/iio.cpp:2754-2776 -- and PFM for 1-channel data;
//           iio_save_image_float (src/iio.h:239, used by src/tvl1occflow_main.cpp:256 for the occlusion
//           map): 8-bit grey PNG (stored deflate blocks, no libpng), PGM or PFM by file extension
//
// The pixel data returned by the readers is malloc'd: the CLI releases it with free()
// (src/tvl1flow_main.cpp:218-219).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool skip_ws_and_comments(FILE *f)
{
    int c;
    while ((c = fgetc(f)) != EOF) {
        if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} continue; }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        ungetc(c, f);
        return true;
    }
    return false;
}

bool read_int(FILE *f, int *v)
{
    return skip_ws_and_comments(f) && fscanf(f, "%d", v) == 1;
}

// grey image as float, row-major; returns false if the format is not understood
bool read_grey(const char *fname, std::vector<float> &out, int *w, int *h)
{
    FILE *f = strcmp(fname, "-") ? fopen(fname, "rb") : stdin;
    if (!f) return false;
    char magic[3] = { 0, 0, 0 };
    bool ok = fread(magic, 1, 2, f) == 2;
    if (ok && magic[0] == 'P' && (magic[1] == '5' || magic[1] == '2')) {
        int maxv = 0;
        ok = read_int(f, w) && read_int(f, h) && read_int(f, &maxv) && *w > 0 && *h > 0 && maxv > 0 && maxv < 65536;
        if (ok) {
            const size_t n = (size_t) *w * *h;
            out.resize(n);
            if (magic[1] == '5') {
                fgetc(f);   // single whitespace after maxval
                if (maxv < 256) {
                    std::vector<unsigned char> b(n);
                    ok = fread(b.data(), 1, n, f) == n;
                    for (size_t i = 0; ok && i < n; i++) out[i] = b[i];
                } else {
                    std::vector<unsigned char> b(2 * n);
                    ok = fread(b.data(), 1, 2 * n, f) == 2 * n;
                    for (size_t i = 0; ok && i < n; i++) out[i] = (float) ((b[2 * i] << 8) | b[2 * i + 1]);
                }
            } else {
                for (size_t i = 0; ok && i < n; i++) { int v; ok = read_int(f, &v); out[i] = (float) v; }
            }
        }
    } else if (ok && magic[0] == 'P' && magic[1] == 'f') {
        float scale = 0.f;
        ok = read_int(f, w) && read_int(f, h) && skip_ws_and_comments(f) && fscanf(f, "%f", &scale) == 1 &&
             *w > 0 && *h > 0;
        if (ok) {
            fgetc(f);
            const size_t n = (size_t) *w * *h;
            std::vector<float> raw(n);
            ok = fread(raw.data(), 4, n, f) == n;
            const uint16_t probe = 1;
            const bool host_little = *(const unsigned char *) &probe == 1;
            if (ok && (scale < 0) != host_little)
                for (float &v : raw) {
                    unsigned char *p = (unsigned char *) &v;
                    std::swap(p[0], p[3]);
                    std::swap(p[1], p[2]);
                }
            out.resize(n);
            for (int y = 0; ok && y < *h; y++)          // PFM stores the bottom row first
                memcpy(&out[(size_t) y * *w], &raw[(size_t) (*h - 1 - y) * *w], sizeof(float) * *w);
        }
    } else {
        ok = false;
    }
    if (f != stdin) fclose(f);
    return ok;
}

template <typename T>
T *read_as(const char *fname, int *w, int *h)
{
    std::vector<float> g;
    if (!read_grey(fname, g, w, h)) {
        fprintf(stderr, "iio_lite: cannot read \"%s\" (supported: PGM P5/P2, PFM Pf)\n", fname);
        return nullptr;
    }
    T *p = (T *) malloc(sizeof(T) * g.size());
    if (!p) return nullptr;
    for (size_t i = 0; i < g.size(); i++) p[i] = (T) g[i];
    return p;
}

} // namespace

float *iio_read_image_float(const char *fname, int *w, int *h) { return read_as<float>(fname, w, h); }

double *iio_read_image_double(const char *fname, int *w, int *h) { return read_as<double>(fname, w, h); }

void iio_save_image_float_vec(const char *filename, float *x, int w, int h, int pd)
{
    FILE *f = strcmp(filename, "-") ? fopen(filename, "wb") : stdout;
    if (!f) { fprintf(stderr, "iio_lite: cannot write \"%s\"\n", filename); exit(EXIT_FAILURE); }
    if (pd == 2) {
        const float pieh = 202021.25f;      // the bytes "PIEH"
        const uint32_t ww = (uint32_t) w, hh = (uint32_t) h;
        fwrite(&pieh, 4, 1, f);
        fwrite(&ww, 4, 1, f);
        fwrite(&hh, 4, 1, f);
        fwrite(x, 4, (size_t) w * h * 2, f);
    } else if (pd == 1) {
        fprintf(f, "Pf\n%d %d\n-1.0\n", w, h);
        for (int y = h - 1; y >= 0; y--) fwrite(x + (size_t) y * w, 4, w, f);
    } else {
        fprintf(stderr, "iio_lite: only 1- or 2-channel float output is supported\n");
        exit(EXIT_FAILURE);
    }
    if (f != stdout) fclose(f);
}

namespace {

uint32_t crc32_of(const unsigned char *p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    if (!table[1])
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

void put_be32(std::vector<unsigned char> &v, uint32_t x)
{
    for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char) (x >> s));
}

void png_chunk(FILE *f, const char *type, const std::vector<unsigned char> &data)
{
    std::vector<unsigned char> b;
    put_be32(b, (uint32_t) data.size());
    b.insert(b.end(), type, type + 4);
    b.insert(b.end(), data.begin(), data.end());
    put_be32(b, crc32_of(b.data() + 4, b.size() - 4));
    fwrite(b.data(), 1, b.size(), f);
}

// 8-bit greyscale PNG whose zlib stream consists of stored (uncompressed) deflate blocks
void write_png_grey8(FILE *f, const std::vector<unsigned char> &px, int w, int h)
{
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t) w);
    put_be32(ihdr, (uint32_t) h);
    const unsigned char tail[5] = { 8, 0, 0, 0, 0 };     // bit depth 8, grey, deflate, no filter, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    png_chunk(f, "IHDR", ihdr);
    std::vector<unsigned char> raw;
    raw.reserve((size_t) h * (w + 1));
    for (int y = 0; y < h; y++) {
        raw.push_back(0);                                  // filter type "none"
        raw.insert(raw.end(), px.begin() + (size_t) y * w, px.begin() + (size_t) (y + 1) * w);
    }
    std::vector<unsigned char> z = { 0x78, 0x01 };
    uint32_t a = 1, b = 0;                                 // Adler-32
    for (unsigned char c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back((unsigned char) (n & 0xFF));
        z.push_back((unsigned char) (n >> 8));
        z.push_back((unsigned char) (~n & 0xFF));
        z.push_back((unsigned char) ((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        if (raw.empty()) break;
    }
    put_be32(z, (b << 16) | a);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
}

bool ends_with(const char *s, const char *ext)
{
    const size_t n = strlen(s), m = strlen(ext);
    return n >= m && !strcmp(s + n - m, ext);
}

} // namespace

void iio_save_image_float(const char *filename, float *x, int w, int h)
{
    if (ends_with(filename, ".pfm")) { iio_save_image_float_vec(filename, x, w, h, 1); return; }
    FILE *f = strcmp(filename, "-") ? fopen(filename, "wb") : stdout;
    if (!f) { fprintf(stderr, "iio_lite: cannot write \"%s\"\n", filename); exit(EXIT_FAILURE); }
    std::vector<unsigned char> px((size_t) w * h);
    for (size_t i = 0; i < px.size(); i++) {
        const float v = x[i] < 0 ? 0 : (x[i] > 255 ? 255 : x[i]);
        px[i] = (unsigned char) (v + 0.5f);
    }
    if (ends_with(filename, ".pgm")) {
        fprintf(f, "P5\n%d %d\n255\n", w, h);
        fwrite(px.data(), 1, px.size(), f);
    } else {
        write_png_grey8(f, px, w, h);
    }
    if (f != stdout) fclose(f);
}
