// tvl1flow_seq -- video / batch front end over the C ABI (include/tvl1_b200.h).
//
// The reference's program (src/tvl1flow_main.cpp) takes exactly two images.  A video is that program
// looped over consecutive frames; this front end does the loop in one process and one library call
// (tvl1_solve_sequence_f32: each frame crosses PCIe once, pairs run as lock-step batches), keeping the
// reference CLI's conventions:
//   * same parameters, defaults and "out of range -> default (+ warning when verbose)" rule
//     (src/tvl1flow_main.cpp:24-33, :112-167);
//   * same nscales clamp from the image diagonal (:185-188);
//   * output: one Middlebury .flo per pair (src/iio.cpp:2754-2776), interleaved (u, v) like :209-214.
//
//   tvl1flow_seq [options] frame0 frame1 [frame2 ...]
//     -o PREFIX   output files PREFIX0000.flo, PREFIX0001.flo, ...   (default "flow_")
//     -p          inputs are independent pairs (f0 f1)(f2 f3)... instead of a sliding sequence
//     -d DEVICE   CUDA device ordinal (default 0)       -b PAIRS  lock-step batch size (default 32)
//     -t tau  -l lambda  -T theta  -s nscales  -z zfactor  -w nwarps  -e epsilon  -v
// Exit status: 0 ok, 1 bad usage / size mismatch / solver failure (message on stderr).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/tvl1_b200.h"

float *iio_read_image_float(const char *fname, int *w, int *h);                       // cli/iio_lite.cpp
void iio_save_image_float_vec(const char *filename, float *x, int w, int h, int pd);

namespace {

int usage(const char *argv0)
{
    fprintf(stderr,
            "usage: %s [-o prefix] [-p] [-d device] [-b batch] [-t tau] [-l lambda] [-T theta]\n"
            "       [-s nscales] [-z zfactor] [-w nwarps] [-e epsilon] [-v] frame0 frame1 [frame2 ...]\n",
            argv0);
    return EXIT_FAILURE;
}

template <typename V>
void fall_back(bool bad, V &value, V dflt, const char *name, const char *fmt, bool verbose)
{
    if (!bad) return;
    value = dflt;
    if (verbose) {
        fprintf(stderr, "warning: %s changed to ", name);
        fprintf(stderr, fmt, value);
        fputc('\n', stderr);
    }
}

} // namespace

int main(int argc, char **argv)
{
    tvl1_params p;
    tvl1_default_params(&p);
    const tvl1_params dflt = p;
    std::string prefix = "flow_";
    bool verbose = false, pairs = false;
    int device = 0, batch = 32;
    std::vector<const char *> files;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] != '-' || a[1] == 0 || a[2] != 0) { files.push_back(a); continue; }
        if (a[1] == 'v') { verbose = true; continue; }
        if (a[1] == 'p') { pairs = true; continue; }
        if (i + 1 >= argc) return usage(argv[0]);
        const char *val = argv[++i];
        switch (a[1]) {
        case 'o': prefix = val; break;
        case 'd': device = atoi(val); break;
        case 'b': batch = atoi(val); break;
        case 't': p.tau = atof(val); break;
        case 'l': p.lambda = atof(val); break;
        case 'T': p.theta = atof(val); break;
        case 's': p.nscales = atoi(val); break;
        case 'z': p.zfactor = atof(val); break;
        case 'w': p.warps = atoi(val); break;
        case 'e': p.epsilon = atof(val); break;
        default: return usage(argv[0]);
        }
    }
    if (files.size() < 2 || (pairs && files.size() % 2)) return usage(argv[0]);

    fall_back(p.tau <= 0 || p.tau > 0.25, p.tau, dflt.tau, "tau", "%g", verbose);
    fall_back(p.lambda <= 0, p.lambda, dflt.lambda, "lambda", "%g", verbose);
    fall_back(p.theta <= 0, p.theta, dflt.theta, "theta", "%g", verbose);
    fall_back(p.nscales <= 0, p.nscales, dflt.nscales, "nscales", "%d", verbose);
    fall_back(p.zfactor <= 0 || p.zfactor >= 1, p.zfactor, dflt.zfactor, "zfactor", "%g", verbose);
    fall_back(p.warps <= 0, p.warps, dflt.warps, "nwarps", "%d", verbose);
    fall_back(p.epsilon <= 0, p.epsilon, dflt.epsilon, "epsilon", "%f", verbose);
    if (batch < 1) batch = 32;

    // frames, all of one size
    int nx = 0, ny = 0;
    std::vector<float> frames;
    for (size_t k = 0; k < files.size(); k++) {
        int w = 0, h = 0;
        float *img = iio_read_image_float(files[k], &w, &h);
        if (!img || w < 1 || h < 1) {
            // iio_lite reads binary/ASCII PGM (P5/P2) and PFM only; the reference's iio also takes PNG/JPEG/TIFF
            fprintf(stderr, "ERROR: cannot read image '%s' (supported: PGM P2/P5, PFM)\n", files[k]);
            free(img);
            return EXIT_FAILURE;
        }
        if (k == 0) { nx = w; ny = h; frames.resize(files.size() * (size_t) nx * ny); }
        if (w != nx || h != ny) {
            fprintf(stderr, "ERROR: input images size mismatch %dx%d != %dx%d\n", nx, ny, w, h);
            return EXIT_FAILURE;
        }
        memcpy(frames.data() + k * (size_t) nx * ny, img, sizeof(float) * nx * ny);
        free(img);
    }
    const size_t n = (size_t) nx * ny;

    // coarsest level no smaller than about 16x16
    const double nmax = 1 + std::log(std::hypot((double) nx, (double) ny) / 16.0) / std::log(1 / p.zfactor);
    if (nmax < p.nscales) p.nscales = (int) nmax;
    if (verbose)
        fprintf(stderr, "frames=%zu tau=%f lambda=%f theta=%f nscales=%d zfactor=%f nwarps=%d epsilon=%g\n",
                files.size(), p.tau, p.lambda, p.theta, p.nscales, p.zfactor, p.warps, p.epsilon);

    tvl1_ctx *ctx = nullptr;
    if (tvl1_create(device, &ctx) != TVL1_OK) {
        fprintf(stderr, "ERROR: %s\n", tvl1_last_error(nullptr));
        return EXIT_FAILURE;
    }
    tvl1_set_max_batch(ctx, batch);

    const int npairs = pairs ? (int) files.size() / 2 : (int) files.size() - 1;
    const int nstat = p.nscales * p.warps;
    std::vector<float> u((size_t) npairs * n), v((size_t) npairs * n);
    std::vector<int> iters((size_t) npairs * nstat);
    std::vector<double> errs((size_t) npairs * nstat);
    int rc;
    if (pairs) {
        // de-interleave (f0 f1)(f2 f3)... into the two planes-of-pairs the batch call takes
        std::vector<float> a((size_t) npairs * n), b((size_t) npairs * n);
        for (int k = 0; k < npairs; k++) {
            memcpy(a.data() + k * n, frames.data() + (2 * (size_t) k) * n, sizeof(float) * n);
            memcpy(b.data() + k * n, frames.data() + (2 * (size_t) k + 1) * n, sizeof(float) * n);
        }
        rc = tvl1_solve_batch_f32(ctx, npairs, a.data(), b.data(), u.data(), v.data(), nx, ny, &p,
                                  iters.data(), errs.data());
    } else {
        rc = tvl1_solve_sequence_f32(ctx, (int) files.size(), frames.data(), u.data(), v.data(), nx, ny, &p,
                                     iters.data(), errs.data());
    }
    if (rc != TVL1_OK) {
        fprintf(stderr, "ERROR: %s\n", rc == TVL1_ERR_SIGMA ? "GaussianSmooth: sigma too large" : tvl1_last_error(ctx));
        tvl1_destroy(ctx);
        return EXIT_FAILURE;
    }

    std::vector<int> sx(p.nscales), sy(p.nscales);
    sx[0] = nx; sy[0] = ny;
    for (int s = 1; s < p.nscales; s++) tvl1_zoom_size(sx[s - 1], sy[s - 1], &sx[s], &sy[s], p.zfactor);
    std::vector<float> f(2 * n);
    for (int k = 0; k < npairs; k++) {
        if (verbose) {
            fprintf(stderr, "Pair %d\n", k);
            for (int s = p.nscales - 1; s >= 0; s--) {
                fprintf(stderr, "Scale %d: %dx%d\n", s, sx[s], sy[s]);
                for (int w = 0; w < p.warps; w++) {
                    const size_t q = (size_t) k * nstat + (size_t) (p.nscales - 1 - s) * p.warps + w;
                    fprintf(stderr, "Warping: %d, Iterations: %d, Error: %f\n", w, iters[q], errs[q]);
                }
            }
        }
        for (size_t i = 0; i < n; i++) {
            f[2 * i] = u[k * n + i];
            f[2 * i + 1] = v[k * n + i];
        }
        char name[32];
        snprintf(name, sizeof name, "%04d.flo", k);
        iio_save_image_float_vec((prefix + name).c_str(), f.data(), nx, ny, 2);
    }
    tvl1_destroy(ctx);
    return EXIT_SUCCESS;
}
