// TEST INFRASTRUCTURE -- never used by the product path.
// C entry points into the post-processing functions the reference keeps in its command-line translation unit
// (leftright_test mgm.cc:68-91, update_dmin_dmax mgm.cc:120-158) and in img_tools.h (median_filter :203-238).
// The reference file is compiled unmodified from where it lies (include path from the Makefile) with its main()
// renamed, so that these functions can be called on arrays.  Nothing here does image I/O: the two iio symbols the
// reference links against are stubs.
#include <stdlib.h>
#include <string.h>

#define main mgm_reference_cli_main
#include "mgm.cc"
#undef main

extern "C" float *iio_read_image_float_split(const char *, int *, int *, int *) { abort(); }
extern "C" void iio_save_image_float_split(char *, float *, int, int, int) { abort(); }

static Img img_from(const float *p, int nx, int ny, int nch) {
   Img r(nx, ny, nch);
   memcpy(&r.data[0], p, sizeof(float) * (size_t)nx * ny * nch);
   return r;
}

extern "C" {

void refcli_leftright(float *dx, int nx, int ny, const float *Rdx, int rnx, int rny, float threshold) {
   Img a = img_from(dx, nx, ny, 1), b = img_from(Rdx, rnx, rny, 1);
   leftright_test(a, b, threshold);
   memcpy(dx, &a.data[0], sizeof(float) * (size_t)nx * ny);
}

void refcli_median(const float *u, int nx, int ny, int nch, int radius, float *out) {
   Img a = img_from(u, nx, ny, nch);
   Img m = median_filter(a, radius);
   memcpy(out, &m.data[0], sizeof(float) * (size_t)nx * ny * nch);
}

void refcli_update_dmin_dmax(const float *off, int nx, int ny, float *lo, float *hi, int slack, int radius, float *mm) {
   Img o = img_from(off, nx, ny, 1), l = img_from(lo, nx, ny, 1), h = img_from(hi, nx, ny, 1);
   std::pair<float, float> g = update_dmin_dmax(o, &l, &h, slack, radius);
   memcpy(lo, &l.data[0], sizeof(float) * (size_t)nx * ny);
   memcpy(hi, &h.data[0], sizeof(float) * (size_t)nx * ny);
   mm[0] = g.first;
   mm[1] = g.second;
}

}  // extern "C"
