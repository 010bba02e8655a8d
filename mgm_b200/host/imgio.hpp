// Minimal image I/O for the mgm CLI (replaces the role of iio/iio.c, which is third-party and needs
// libpng/libtiff/libjpeg).  Reads PNM (P2/P3/P5/P6, 8/16 bit), PFM (Pf/PF), NPY (u1/u2/f4/f8, C order),
// PNG (8/16 bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced, via zlib) and uncompressed TIFF
// (8/16-bit integer, 32-bit float, chunky, strips).  Writes NPY, PFM and uncompressed float32 TIFF, chosen by
// the file extension.  Images are returned planar: data[x + y*nx + c*nx*ny] (iio_read_image_float_split).
// Row order follows iio: rows are stored top-down in every format, including PFM (iio.c:2561-2579,4112-4125).
// Every offset and length taken from a file is validated against the file size before it is used: a truncated or
// malformed input ends in a runtime_error (the CLI prints it and exits with code 3), never in an out-of-bounds read.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "mgmb200_host.hpp"

namespace mgmb200 {
namespace io {

inline std::vector<uint8_t> slurp(const char *path) {
   FILE *f = fopen(path, "rb");
   if (!f) throw std::runtime_error(std::string("cannot open ") + path);
   std::vector<uint8_t> b;
   uint8_t tmp[1 << 16];
   size_t n;
   while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) b.insert(b.end(), tmp, tmp + n);
   fclose(f);
   return b;
}

// [off, off+len) must lie inside a buffer of `size` bytes
inline void need(size_t size, size_t off, size_t len, const char *what) {
   if (off > size || len > size - off) throw std::runtime_error(std::string("truncated or malformed ") + what);
}
// image dimensions from a header: positive, and small enough that nx*ny*nch floats can be indexed and allocated
inline void check_dims(long long nx, long long ny, long long nch, const char *what) {
   if (nx < 1 || ny < 1 || nch < 1 || nx > (1 << 24) || ny > (1 << 24) || nch > 64 || nx * ny * nch > (1LL << 33))
      throw std::runtime_error(std::string("implausible image size in ") + what);
}

// interleaved (x,y,c) -> planar Img
inline Img from_interleaved(const std::vector<float> &v, int nx, int ny, int nch) {
   Img im(nx, ny, nch);
   for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++)
         for (int c = 0; c < nch; c++) im.data[x + (size_t)y * nx + (size_t)c * nx * ny] = v[((size_t)y * nx + x) * nch + c];
   return im;
}

// ---------------------------------------------------------------- PNM / PFM
inline int pnm_token(const std::vector<uint8_t> &b, size_t &p, std::string &tok) {
   tok.clear();
   while (p < b.size()) {
      if (b[p] == '#') { while (p < b.size() && b[p] != '\n') p++; }
      else if (isspace(b[p])) p++;
      else break;
   }
   while (p < b.size() && !isspace(b[p])) tok.push_back((char)b[p++]);
   return !tok.empty();
}

inline Img read_pnm(const std::vector<uint8_t> &b) {
   size_t p = 0;
   std::string t;
   pnm_token(b, p, t);
   const std::string magic = t;
   if (magic == "Pf" || magic == "PF") {
      const int nch = magic == "PF" ? 3 : 1;
      pnm_token(b, p, t); const int nx = atoi(t.c_str());
      pnm_token(b, p, t); const int ny = atoi(t.c_str());
      pnm_token(b, p, t); const double scale = atof(t.c_str());
      p++;   // single whitespace after the scale
      check_dims(nx, ny, nch, "PFM header");
      need(b.size(), p, (size_t)nx * ny * nch * 4, "PFM");
      std::vector<float> v((size_t)nx * ny * nch);
      memcpy(v.data(), &b[p], v.size() * 4);
      if (scale > 0)   // big endian payload
         for (float &f : v) { uint32_t u; memcpy(&u, &f, 4); u = __builtin_bswap32(u); memcpy(&f, &u, 4); }
      return from_interleaved(v, nx, ny, nch);
   }
   if (magic.size() != 2 || magic[0] != 'P') throw std::runtime_error("not a PNM file");
   const int kind = magic[1] - '0';
   if (kind != 2 && kind != 3 && kind != 5 && kind != 6) throw std::runtime_error("unsupported PNM kind " + magic);
   const int nch = (kind == 3 || kind == 6) ? 3 : 1;
   pnm_token(b, p, t); const int nx = atoi(t.c_str());
   pnm_token(b, p, t); const int ny = atoi(t.c_str());
   pnm_token(b, p, t); const int maxval = atoi(t.c_str());
   check_dims(nx, ny, nch, "PNM header");
   if (maxval < 1 || maxval > 65535) throw std::runtime_error("bad PNM maxval");
   if (kind == 2 || kind == 3) need(b.size(), p, (size_t)nx * ny * nch, "PNM");   // at least one byte per sample
   std::vector<float> v((size_t)nx * ny * nch);
   if (kind == 2 || kind == 3) {
      for (float &f : v) { if (!pnm_token(b, p, t)) throw std::runtime_error("truncated PNM"); f = (float)atoi(t.c_str()); }
   } else {
      p++;
      const int bps = maxval > 255 ? 2 : 1;
      need(b.size(), p, v.size() * bps, "PNM");
      for (size_t i = 0; i < v.size(); i++)
         v[i] = bps == 1 ? (float)b[p + i] : (float)((b[p + 2 * i] << 8) | b[p + 2 * i + 1]);
   }
   return from_interleaved(v, nx, ny, nch);
}

// ---------------------------------------------------------------- NPY
inline Img read_npy(const std::vector<uint8_t> &b) {
   if (b.size() < 12 || memcmp(b.data(), "\x93NUMPY", 6)) throw std::runtime_error("not an NPY file");
   const int major = b[6];
   size_t hlen = major == 1 ? (b[8] | (b[9] << 8)) : (b[8] | (b[9] << 8) | (b[10] << 16) | ((size_t)b[11] << 24));
   const size_t hoff = major == 1 ? 10 : 12;
   need(b.size(), hoff, hlen, "NPY header");
   const std::string h((const char *)&b[hoff], hlen);
   auto field = [&](const char *key) { size_t k = h.find(key); if (k == std::string::npos) throw std::runtime_error("bad NPY header"); return k; };
   size_t k = field("'descr'");
   k = h.find('\'', k + 7);
   if (k == std::string::npos || h.find('\'', k + 1) == std::string::npos) throw std::runtime_error("bad NPY header");
   const std::string descr = h.substr(k + 1, h.find('\'', k + 1) - k - 1);
   if (h.find("'fortran_order': True") != std::string::npos) throw std::runtime_error("fortran-order NPY not supported");
   k = field("'shape'");
   k = h.find('(', k);
   if (k == std::string::npos || h.find(')', k) == std::string::npos) throw std::runtime_error("bad NPY header");
   std::vector<int> shape;
   const std::string sh = h.substr(k + 1, h.find(')', k) - k - 1);
   for (size_t i = 0; i < sh.size();) {
      while (i < sh.size() && !isdigit(sh[i])) i++;
      if (i >= sh.size()) break;
      shape.push_back(atoi(sh.c_str() + i));
      while (i < sh.size() && isdigit(sh[i])) i++;
   }
   if (shape.size() < 2 || shape.size() > 3) throw std::runtime_error("NPY image must be (H,W) or (H,W,C)");
   const int ny = shape[0], nx = shape[1], nch = shape.size() == 3 ? shape[2] : 1;
   check_dims(nx, ny, nch, "NPY header");
   const size_t esz = (descr == "<f4" || descr == "<i4") ? 4 : descr == "<f8" ? 8 : descr == "|u1" ? 1 : descr == "<u2" ? 2 : 0;
   if (!esz) throw std::runtime_error("unsupported NPY dtype " + descr);
   need(b.size(), hoff + hlen, (size_t)nx * ny * nch * esz, "NPY payload");
   const uint8_t *d = b.data() + hoff + hlen;
   std::vector<float> v((size_t)nx * ny * nch);
   for (size_t i = 0; i < v.size(); i++) {
      if (descr == "<f4") { memcpy(&v[i], d + 4 * i, 4); }
      else if (descr == "<f8") { double t; memcpy(&t, d + 8 * i, 8); v[i] = (float)t; }
      else if (descr == "|u1") v[i] = d[i];
      else if (descr == "<u2") { uint16_t t; memcpy(&t, d + 2 * i, 2); v[i] = t; }
      else if (descr == "<i4") { int32_t t; memcpy(&t, d + 4 * i, 4); v[i] = (float)t; }
      else throw std::runtime_error("unsupported NPY dtype " + descr);
   }
   return from_interleaved(v, nx, ny, nch);
}

// ---------------------------------------------------------------- PNG (zlib inflate + scanline filters)
inline Img read_png(const std::vector<uint8_t> &b) {
   auto be32 = [&](size_t p) { return ((uint32_t)b[p] << 24) | (b[p + 1] << 16) | (b[p + 2] << 8) | b[p + 3]; };
   size_t p = 8;
   uint32_t w = 0, h = 0;
   int depth = 0, ctype = 0, interlace = 0;
   std::vector<uint8_t> idat, plte;
   while (p + 8 <= b.size()) {
      const uint32_t len = be32(p);
      const std::string type((const char *)&b[p + 4], 4);
      const size_t d = p + 8;
      need(b.size(), d, (size_t)len + 4, "PNG chunk");   // data + CRC
      if (type == "IHDR") {
         if (len < 13) throw std::runtime_error("bad PNG header");
         w = be32(d); h = be32(d + 4); depth = b[d + 8]; ctype = b[d + 9]; interlace = b[d + 12];
      }
      else if (type == "PLTE") plte.assign(b.begin() + d, b.begin() + d + len);
      else if (type == "IDAT") idat.insert(idat.end(), b.begin() + d, b.begin() + d + len);
      else if (type == "IEND") break;
      p = d + len + 4;
   }
   if (interlace) throw std::runtime_error("interlaced PNG not supported");
   if (depth != 8 && depth != 16) throw std::runtime_error("PNG bit depth must be 8 or 16");
   if (ctype != 0 && ctype != 2 && ctype != 3 && ctype != 4 && ctype != 6) throw std::runtime_error("bad PNG colour type");
   if (ctype == 3 && depth != 8) throw std::runtime_error("palette PNG must be 8 bit");
   check_dims(w, h, 4, "PNG header");
   const int chans = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
   const int bpp = chans * depth / 8;
   const size_t stride = (size_t)w * bpp;
   std::vector<uint8_t> raw((stride + 1) * h);
   uLongf rawlen = raw.size();
   if (idat.empty() || uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK) throw std::runtime_error("PNG inflate failed");
   if (rawlen != raw.size()) throw std::runtime_error("PNG pixel data shorter than the header says");
   std::vector<uint8_t> img(stride * h);
   for (uint32_t y = 0; y < h; y++) {
      const uint8_t *src = &raw[(stride + 1) * y];
      uint8_t *dst = &img[stride * y];
      const uint8_t *up = y ? dst - stride : nullptr;
      const int ft = src[0];
      for (size_t i = 0; i < stride; i++) {
         const int a = i >= (size_t)bpp ? dst[i - bpp] : 0, bb = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
         int pred = 0;
         if (ft == 1) pred = a;
         else if (ft == 2) pred = bb;
         else if (ft == 3) pred = (a + bb) >> 1;
         else if (ft == 4) { const int pp = a + bb - c, pa = abs(pp - a), pb = abs(pp - bb), pc = abs(pp - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c); }
         dst[i] = (uint8_t)(src[i + 1] + pred);
      }
   }
   // colour channels only (alpha dropped like a plain RGB read), palette expanded
   const int outch = (ctype == 2 || ctype == 6 || ctype == 3) ? 3 : 1;
   std::vector<float> v((size_t)w * h * outch);
   for (size_t i = 0; i < (size_t)w * h; i++)
      for (int c = 0; c < outch; c++) {
         if (ctype == 3) v[i * outch + c] = plte.at(3 * img[i] + c);
         else {
            const size_t o = i * bpp + (size_t)c * depth / 8;
            v[i * outch + c] = depth == 8 ? (float)img[o] : (float)((img[o] << 8) | img[o + 1]);
         }
      }
   return from_interleaved(v, (int)w, (int)h, outch);
}

// ---------------------------------------------------------------- TIFF (uncompressed strips)
inline Img read_tiff(const std::vector<uint8_t> &b) {
   const bool le = b[0] == 'I';
   auto u16 = [&](size_t p) -> uint32_t { need(b.size(), p, 2, "TIFF"); return le ? (b[p] | (b[p + 1] << 8)) : ((b[p] << 8) | b[p + 1]); };
   auto u32 = [&](size_t p) -> uint32_t { need(b.size(), p, 4, "TIFF");
                                          return le ? (b[p] | (b[p + 1] << 8) | (b[p + 2] << 16) | ((uint32_t)b[p + 3] << 24))
                                                    : (((uint32_t)b[p] << 24) | (b[p + 1] << 16) | (b[p + 2] << 8) | b[p + 3]); };
   if (u16(2) != 42) throw std::runtime_error("not a classic TIFF");
   size_t ifd = u32(4);
   const int n = u16(ifd);
   need(b.size(), ifd + 2, (size_t)12 * n, "TIFF directory");
   uint32_t w = 0, h = 0, bps = 8, comp = 1, spp = 1, fmt = 1, planar = 1;
   std::vector<uint32_t> offs, counts;
   for (int i = 0; i < n; i++) {
      const size_t e = ifd + 2 + 12 * i;
      const uint32_t tag = u16(e), type = u16(e + 2), cnt = u32(e + 4);
      const size_t tsz = type == 3 ? 2 : 4;
      if ((tag == 273 || tag == 279) && (type != 3 && type != 4)) throw std::runtime_error("unsupported TIFF strip table type");
      if ((tag == 273 || tag == 279) && cnt > b.size()) throw std::runtime_error("malformed TIFF strip table");
      auto value = [&](uint32_t j) -> uint32_t {
         const size_t base = (cnt * tsz <= 4) ? e + 8 : u32(e + 8);
         return type == 3 ? u16(base + 2 * j) : u32(base + 4 * j);
      };
      if (tag == 256) w = value(0); else if (tag == 257) h = value(0); else if (tag == 258) bps = value(0);
      else if (tag == 259) comp = value(0); else if (tag == 277) spp = value(0); else if (tag == 339) fmt = value(0);
      else if (tag == 284) planar = value(0);   // rows per strip (278) is implied by the strip offsets
      else if (tag == 273) { offs.resize(cnt); for (uint32_t j = 0; j < cnt; j++) offs[j] = value(j); }
      else if (tag == 279) { counts.resize(cnt); for (uint32_t j = 0; j < cnt; j++) counts[j] = value(j); }
   }
   if (comp != 1) throw std::runtime_error("compressed TIFF is not supported by this reader (convert to uncompressed TIFF, PFM or NPY)");
   if (planar != 1 && spp > 1) throw std::runtime_error("planar TIFF not supported");
   check_dims(w, h, spp, "TIFF directory");
   if (offs.size() != counts.size()) throw std::runtime_error("malformed TIFF strip table");
   std::vector<uint8_t> pix;
   for (size_t s = 0; s < offs.size(); s++) {
      need(b.size(), offs[s], counts[s], "TIFF strip");
      pix.insert(pix.end(), b.begin() + offs[s], b.begin() + offs[s] + counts[s]);
   }
   std::vector<float> v((size_t)w * h * spp);
   if (bps != 8 && bps != 16 && bps != 32) throw std::runtime_error("unsupported TIFF sample size");
   const size_t by = bps / 8;
   if (pix.size() < v.size() * by) throw std::runtime_error("truncated TIFF");
   for (size_t i = 0; i < v.size(); i++) {
      const uint8_t *d = &pix[i * by];
      if (bps == 8) v[i] = d[0];
      else if (bps == 16) v[i] = (float)(le ? (d[0] | (d[1] << 8)) : ((d[0] << 8) | d[1]));
      else if (bps == 32 && fmt == 3) { uint32_t t = le ? (d[0] | (d[1] << 8) | (d[2] << 16) | ((uint32_t)d[3] << 24)) : (((uint32_t)d[0] << 24) | (d[1] << 16) | (d[2] << 8) | d[3]); memcpy(&v[i], &t, 4); }
      else throw std::runtime_error("unsupported TIFF sample format");
   }
   return from_interleaved(v, (int)w, (int)h, (int)spp);
}

inline Img read_image(const char *path) {
   const std::vector<uint8_t> b = slurp(path);
   if (b.size() < 8) throw std::runtime_error(std::string("empty or truncated image file: ") + path);
   if (b.size() >= 8 && !memcmp(b.data(), "\x89PNG\r\n\x1a\n", 8)) return read_png(b);
   if (b.size() >= 6 && !memcmp(b.data(), "\x93NUMPY", 6)) return read_npy(b);
   if (b.size() >= 4 && ((b[0] == 'I' && b[1] == 'I') || (b[0] == 'M' && b[1] == 'M'))) return read_tiff(b);
   if (b.size() >= 2 && b[0] == 'P') return read_pnm(b);
   throw std::runtime_error(std::string("unrecognised image format: ") + path);
}

// ---------------------------------------------------------------- writers (float32)
inline std::vector<float> to_interleaved(const Img &im) {
   std::vector<float> v(im.data.size());
   for (int y = 0; y < im.ny; y++)
      for (int x = 0; x < im.nx; x++)
         for (int c = 0; c < im.nch; c++) v[((size_t)y * im.nx + x) * im.nch + c] = im.data[x + (size_t)y * im.nx + (size_t)c * im.nx * im.ny];
   return v;
}

inline void write_image(const char *path, const Img &im) {
   const std::string p(path);
   const std::string ext = p.rfind('.') == std::string::npos ? "" : p.substr(p.rfind('.') + 1);
   const std::vector<float> v = to_interleaved(im);
   // formats are chosen by the extension like iio does; the ones this build cannot encode are refused instead of
   // writing other bytes under that name (no extension = PFM, what the reference built without image libraries writes)
   const bool pfm = ext == "pfm" || ext == "PFM" || ext.empty() || p.rfind('.') < p.rfind('/') + 1;
   if (!(ext == "npy" || ext == "tif" || ext == "tiff" || pfm))
      throw std::runtime_error("cannot write " + p + ": supported output formats are .npy, .tif/.tiff (float32) and .pfm");
   if (pfm && im.nch != 1 && im.nch != 3) throw std::runtime_error("cannot write " + p + ": PFM holds 1 or 3 channels, use .npy or .tif");
   FILE *f = fopen(path, "wb");
   if (!f) throw std::runtime_error("cannot write " + p);
   if (ext == "npy") {
      char hdr[256];
      int n = snprintf(hdr, sizeof hdr, "{'descr': '<f4', 'fortran_order': False, 'shape': (%d, %d, %d), }", im.ny, im.nx, im.nch);
      const int total = ((10 + n + 1 + 63) / 64) * 64;
      const uint16_t hlen = (uint16_t)(total - 10);
      fwrite("\x93NUMPY\x01\x00", 1, 8, f);
      fwrite(&hlen, 2, 1, f);
      fwrite(hdr, 1, n, f);
      for (int i = 10 + n; i < total - 1; i++) fputc(' ', f);
      fputc('\n', f);
      fwrite(v.data(), 4, v.size(), f);
   } else if (ext == "tif" || ext == "tiff") {
      // little-endian classic TIFF, one strip, 32-bit IEEE float, chunky
      const uint32_t nbytes = (uint32_t)(v.size() * 4), ifd = 8 + nbytes;
      fwrite("II*\0", 1, 4, f);
      fwrite(&ifd, 4, 1, f);
      fwrite(v.data(), 4, v.size(), f);
      struct E { uint16_t tag, type; uint32_t cnt, val; };
      const E es[] = {{256, 4, 1, (uint32_t)im.nx}, {257, 4, 1, (uint32_t)im.ny}, {258, 3, 1, 32}, {259, 3, 1, 1},
                      {262, 3, 1, (uint32_t)(im.nch >= 3 ? 2 : 1)}, {273, 4, 1, 8}, {277, 3, 1, (uint32_t)im.nch},
                      {278, 4, 1, (uint32_t)im.ny}, {279, 4, 1, nbytes}, {284, 3, 1, 1}, {339, 3, 1, 3}};
      const uint16_t ne = sizeof(es) / sizeof(es[0]);
      fwrite(&ne, 2, 1, f);
      for (const E &e : es) { fwrite(&e.tag, 2, 1, f); fwrite(&e.type, 2, 1, f); fwrite(&e.cnt, 4, 1, f); fwrite(&e.val, 4, 1, f); }
      const uint32_t zero = 0;
      fwrite(&zero, 4, 1, f);
   } else {
      // PFM the way iio writes it: top-down rows, little endian (negative scale)
      fprintf(f, "%s\n%d %d\n-1\n", im.nch == 3 ? "PF" : "Pf", im.nx, im.ny);
      fwrite(v.data(), 4, v.size(), f);
   }
   fclose(f);
}

}  // namespace io
}  // namespace mgmb200
