// TEST INFRASTRUCTURE.  Drives the reference's own CPU frame cache (src/TNetLib/Cache.cc — the same state machine, leftover
// carry-over, std::random_shuffle + lrand48 permutation and discard rule as the GPU trainer's CuCache, src/CuTNetLib/cuCache.cc)
// over ragged sequences in the trainer's loop order (src/TNet.cc / TNetCu.cc:376-441: AddData until Full, Randomize, GetBunch until
// Empty) and dumps every bunch.  tests/test_oracle_golden.py compares oracle/tnet_oracle.c's cache with this dump bit for bit.
// Only THIS file is ours; everything it calls is the reference.
//
//   cache_dump <in.bin> <out.bin>
// in : int32 cachesize, bunchsize, seed, randomize, fdim, ddim, nseq, then nseq x { int32 rows, float32 rows*fdim, float32 rows*ddim }
// out: int32 nbunches, then per bunch float32 bunch*fdim, float32 bunch*ddim ; finally int32 discarded
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Cache.h"
#include "Matrix.h"

using namespace TNet;

static int get32(FILE *f) { int v = 0; if (fread(&v, 4, 1, f) != 1) { fprintf(stderr, "short input\n"); exit(2); } return v; }

int main(int argc, char **argv) {
  if (argc != 3) { fprintf(stderr, "usage: cache_dump in.bin out.bin\n"); return 2; }
  FILE *in = fopen(argv[1], "rb");
  if (!in) { perror("in"); return 1; }
  const int cachesize = get32(in), bunchsize = get32(in), seed = get32(in), randomize = get32(in), fdim = get32(in), ddim = get32(in),
            nseq = get32(in);
  std::vector<Matrix<BaseFloat> *> F, D;
  for (int s = 0; s < nseq; s++) {
    const int rows = get32(in);
    Matrix<BaseFloat> *f = new Matrix<BaseFloat>(rows, fdim), *d = new Matrix<BaseFloat>(rows, ddim);
    for (int r = 0; r < rows; r++) if (fread(f->pRowData(r), sizeof(float), fdim, in) != (size_t)fdim) return 2;
    for (int r = 0; r < rows; r++) if (fread(d->pRowData(r), sizeof(float), ddim, in) != (size_t)ddim) return 2;
    F.push_back(f); D.push_back(d);
  }
  fclose(in);
  try {
    Cache cache;
    cache.Init(cachesize, bunchsize, seed);
    FILE *out = fopen(argv[2], "wb");
    int nb = 0;
    fwrite(&nb, 4, 1, out);
    Matrix<BaseFloat> feats, labs;
    size_t i = 0;
    while (i < F.size()) {
      while (!cache.Full() && i < F.size()) { cache.AddData(*F[i], *D[i]); i++; }
      if (randomize) cache.Randomize();
      while (!cache.Empty()) {
        cache.GetBunch(feats, labs);
        for (size_t r = 0; r < feats.Rows(); r++) fwrite(feats.pRowData(r), sizeof(float), fdim, out);
        for (size_t r = 0; r < labs.Rows(); r++) fwrite(labs.pRowData(r), sizeof(float), ddim, out);
        nb++;
      }
    }
    const int disc = cache.Discarded();
    fwrite(&disc, 4, 1, out);
    fseek(out, 0, SEEK_SET);
    fwrite(&nb, 4, 1, out);
    fclose(out);
  } catch (std::exception &e) {
    fprintf(stderr, "cache_dump: %s\n", e.what());
    return 1;
  }
  return 0;
}
