// Test helper (tests/test_host_math.py::test_image_reader_rejects_malformed_files): reads an image with the CLI's
// reader and prints its size, or the error.
#include "imgio.hpp"
int main(int argc, char **argv) {
   try {
      mgmb200::Img im = mgmb200::io::read_image(argv[1]);
      if (argc > 2) mgmb200::io::write_image(argv[2], im);
      printf("ok %d %d %d\n", im.nx, im.ny, im.nch);
      return 0;
   } catch (const std::exception &e) {
      printf("error %s\n", e.what());
      return 3;
   }
}
