// Standalone aggregator with the file protocol of matlab/mgm_o.cc:509-609 (MGM_wrapper.m:82-108):
//   mgm_o input.bin labeling.bin [P1 P2 MGM VTYPE]
// input.bin  = int32 ncol, nrow, nlab, NDIR; float32 costs[i + o*ncol*nrow]; float32 w[i + k*ncol*nrow] (8 planes)
// output.bin = ncol*nrow float32 labels, row-major.
#include "mgmb200_host.hpp"

int main(int argc, char *argv[]) {
   if (argc < 3) {
      fprintf(stderr, "too few parameters\n   usage: %s input.bin labeling.bin [P1  P2  MGM  VTYPE]\n", argv[0]);
      return 1;
   }
   int i = 3;
   const float P1 = (argc > i) ? atof(argv[i]) : 8; i++;
   const float P2 = (argc > i) ? atof(argv[i]) : 32; i++;
   const int vMGM = (argc > i) ? atoi(argv[i]) : 2; i++;
   const int felz = (argc > i) ? atoi(argv[i]) : 0; i++;
   FILE *fp = fopen(argv[1], "rb");
   if (!fp) { printf("Error: I cannot open input file\n"); return 1; }
   printf("Reading input data...");
   int hdr[4];
   if (fread(hdr, sizeof(int), 4, fp) != 4) { printf("Error: short header\n"); return 1; }
   const int ncol = hdr[0], nrow = hdr[1], nlab = hdr[2], NDIR = hdr[3];
   std::vector<float> costs((size_t)ncol * nrow * nlab), w((size_t)ncol * nrow * 8), labels((size_t)ncol * nrow);
   if (fread(costs.data(), sizeof(float), costs.size(), fp) != costs.size() ||
       fread(w.data(), sizeof(float), w.size(), fp) != w.size()) { printf("Error: short file\n"); return 1; }
   fclose(fp);
   printf("Done\n");
   printf("#MRF-nodes = %d x %d, #labels = %d, #NDIR = %d, #MGM = %d, P1 = %.2f, P2 = %.2f\n", ncol, nrow, nlab, NDIR, vMGM, P1, P2);
   try {
      mgmb200::check(mgmb200_mgm_labelmajor(mgmb200::context(), costs.data(), w.data(), ncol, nrow, nlab, P1, P2, NDIR, vMGM,
                                            felz, labels.data(), nullptr));
   } catch (const std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 3;
   }
   printf("Writing labels to output file...");
   FILE *outf = fopen(argv[2], "wb");
   fwrite(labels.data(), sizeof(float), labels.size(), outf);
   fclose(outf);
   printf("Done\n");
   return 0;
}
