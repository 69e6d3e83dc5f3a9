/*
 * ort_host.cpp -- host side of the drop-in surface (no CUDA in this file): the positional
 * `*.params` readers, the dispersion laws, the scalar prologue of the reference's main program,
 * the output-name recipe and the raw image / trans-stats writers.
 *
 * Reference behaviour reproduced here: src/setupMod.f90:28-140 (settings.params),
 * src/lens.f90:73-227 (lens / bottle files), src/lens.f90:647-695 (Sellmeier, Cauchy, glass
 * dispersion), src/main.f90:45-70,81 (file name, cosThetaMax, offset guard, annulus radii,
 * image plane), src/imageMod.f90:93-114 (image files), src/main.f90:168-178 (trans-stats.dat),
 * src/utils.f90:351-420 (str()).
 */
#include <sys/stat.h>

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "ort_internal.h"
#include "ort_optics.cuh" /* host build of the counter-based generator (init_emit_image's draws) */

namespace {

/* ------------------------------------------------------------------------------------------
 * Fortran list-directed input, as far as the params files use it: every `read(u,*) x` takes
 * the first value of the next non-blank record and ignores the rest of the line.
 * ---------------------------------------------------------------------------------------- */
class ListReader {
   public:
    explicit ListReader(const std::string& path) : path_(path) {
        std::ifstream in(path);
        ok_ = in.good();
        std::string line;
        while (std::getline(in, line)) {
            std::string tok = first_value(line);
            if (!tok.empty()) vals_.push_back(tok);
        }
    }
    bool ok() const { return ok_; }
    size_t remaining() const { return vals_.size() - next_; }
    const std::string& path() const { return path_; }

    bool real(double* out, const char* what) {
        std::string t;
        if (!take(&t, what)) return false;
        std::string c = t;
        for (char& ch : c)
            if (ch == 'd' || ch == 'D' || ch == 'q' || ch == 'Q') ch = 'e';
        char* end = nullptr;
        errno = 0;
        double v = std::strtod(c.c_str(), &end);
        if (end == c.c_str() || *end != '\0') {
            ort_set_error("%s: line %zu (%s): '%s' is not a real", path_.c_str(), next_, what, t.c_str());
            return false;
        }
        *out = v;
        return true;
    }
    bool integer(int64_t* out, const char* what) {
        std::string t;
        if (!take(&t, what)) return false;
        char* end = nullptr;
        long long v = std::strtoll(t.c_str(), &end, 10);
        if (end == t.c_str() || *end != '\0') {
            ort_set_error("%s: line %zu (%s): '%s' is not an integer", path_.c_str(), next_, what, t.c_str());
            return false;
        }
        *out = v;
        return true;
    }
    bool logical(int32_t* out, const char* what) {
        std::string t;
        if (!take(&t, what)) return false;
        size_t i = (t[0] == '.') ? 1 : 0;
        char ch = i < t.size() ? (char)std::tolower((unsigned char)t[i]) : '?';
        if (ch == 't') {
            *out = 1;
        } else if (ch == 'f') {
            *out = 0;
        } else {
            ort_set_error("%s: line %zu (%s): '%s' is not a logical", path_.c_str(), next_, what, t.c_str());
            return false;
        }
        return true;
    }
    bool text(char* out, size_t cap, const char* what) {
        std::string t;
        if (!take(&t, what)) return false;
        if (t.size() >= cap) {
            ort_set_error("%s: %s too long", path_.c_str(), what);
            return false;
        }
        std::memcpy(out, t.c_str(), t.size() + 1);
        return true;
    }

   private:
    static std::string first_value(const std::string& line) {
        size_t i = 0;
        while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
        if (i >= line.size()) return "";
        if (line[i] == '\'' || line[i] == '"') {
            char qc = line[i];
            size_t j = line.find(qc, i + 1);
            return line.substr(i + 1, j == std::string::npos ? std::string::npos : j - i - 1);
        }
        size_t j = i;
        while (j < line.size() && !std::isspace((unsigned char)line[j]) && line[j] != ',' && line[j] != '/')
            ++j;
        return line.substr(i, j - i);
    }
    bool take(std::string* t, const char* what) {
        if (next_ >= vals_.size()) {
            ort_set_error("%s: end of file while reading %s (value %zu)", path_.c_str(), what, next_ + 1);
            return false;
        }
        *t = vals_[next_++];
        return true;
    }
    std::string path_;
    std::vector<std::string> vals_;
    size_t next_ = 0;
    bool ok_ = false;
};

/* dispersion laws; wavelength arrives in metres and is used in micrometres */
double sellmeier(double wave, const double B[3], const double Cc[3]) { /* src/lens.f90:647-665 */
    double um = wave * 1e6, w2 = um * um, acc = 0.0;
    double term[3];
    for (int i = 0; i < 3; ++i) term[i] = (B[i] * w2) / (w2 - Cc[i]);
    acc = (term[0] + term[1]) + term[2];
    return std::sqrt(1.0 + acc);
}
double cauchy_index(double wave, double a, double b, double c) { /* src/lens.f90:667-680 */
    double um = wave * 1e6, w2 = um * um;
    return a + b * (1.0 / w2) + c * (1.0 / (w2 * w2));
}
double glass_dispersion(double wave, double a, double b, double c) { /* src/lens.f90:682-695 */
    double um = wave * 1e6, w2 = um * um;
    return a - b * w2 + (c / w2);
}

std::string join(const char* dir, const char* name) {
    std::string d = dir ? dir : "";
    if (!d.empty() && d.back() != '/') d += '/';
    return d + name;
}

/* str(x, len): first `len` characters of the left-justified f100.16 rendering
 * (src/utils.f90:351-369) */
std::string str_real(double x, int len) {
    char buf[160];
    std::snprintf(buf, sizeof buf, "%.16f", x);
    std::string s(buf);
    if ((int)s.size() > len) s.resize(len);
    while (!s.empty() && s.back() == ' ') s.pop_back();
    return s;
}

/* gfortran-style list-directed rendering of a real(8) (best effort) */
std::string list_real(double x) {
    char buf[64];
    double ax = std::fabs(x);
    if (x == 0.0) {
        std::snprintf(buf, sizeof buf, "%21.16f    ", x);
    } else if (ax >= 0.1 && ax < 1e16) {
        int intdigits = (ax < 1.0) ? 0 : (int)std::floor(std::log10(ax)) + 1;
        int dec = 17 - (intdigits > 0 ? intdigits : 0);
        if (intdigits == 0) dec = 17;
        std::snprintf(buf, sizeof buf, "%*.*f    ", 21, dec, x);
    } else {
        char e[64];
        std::snprintf(e, sizeof e, "%.16E", x);
        /* C prints E-03, gfortran E-003 */
        std::string es(e);
        size_t p = es.find('E');
        std::string mant = es.substr(0, p), ex = es.substr(p + 2);
        while (ex.size() < 3) ex = "0" + ex;
        std::snprintf(buf, sizeof buf, "%25s", (mant + "E" + es[p + 1] + ex).c_str());
    }
    return buf;
}

}  // namespace

extern "C" {

int ort_struct_sizes(int32_t out[8]) {
    if (!out) return ORT_EINVAL;
    out[0] = (int32_t)sizeof(ort_plano);
    out[1] = (int32_t)sizeof(ort_doublet);
    out[2] = (int32_t)sizeof(ort_bottle);
    out[3] = (int32_t)sizeof(ort_scene);
    out[4] = (int32_t)sizeof(ort_job);
    out[5] = (int32_t)sizeof(ort_timing);
    out[6] = (int32_t)sizeof(ort_settings);
    out[7] = ORT_VERSION;
    return ORT_OK;
}

/* init_plano_convex, src/lens.f90:129-167 */
int ort_load_plano(const char* path, double wavelength, double offset, ort_plano* out) {
    if (!path || !out) return ORT_EINVAL;
    ListReader rd(path);
    if (!rd.ok()) {
        ort_set_error("cannot open %s", path);
        return ORT_EIO;
    }
    double B[3], Cc[3];
    std::memset(out, 0, sizeof *out);
    if (!rd.real(&out->thickness, "thickness") || !rd.real(&out->curve_radius, "curve_radius") ||
        !rd.real(&out->diameter, "diameter") || !rd.real(&out->f, "f") || !rd.real(&out->fb, "fb") ||
        !rd.real(&out->n1, "n1") || !rd.real(&B[0], "b1") || !rd.real(&B[1], "b2") || !rd.real(&B[2], "b3") ||
        !rd.real(&Cc[0], "c1") || !rd.real(&Cc[1], "c2") || !rd.real(&Cc[2], "c3"))
        return ORT_EPARSE;
    out->n2 = sellmeier(wavelength, B, Cc);
    out->radius = out->diameter / 2.0;
    out->centre[2] = offset + (out->fb + out->thickness) - out->curve_radius;
    out->flat_normal[2] = -1.0;
    return ORT_OK;
}

/* init_achromatic_doublet, src/lens.f90:73-126 */
int ort_load_doublet(const char* path, double wavelength, double offset, ort_doublet* out) {
    if (!path || !out) return ORT_EINVAL;
    ListReader rd(path);
    if (!rd.ok()) {
        ort_set_error("cannot open %s", path);
        return ORT_EIO;
    }
    double B1[3], C1[3], B2[3], C2[3];
    std::memset(out, 0, sizeof *out);
    bool good = rd.real(&out->thickness1, "thickness1") && rd.real(&out->thickness2, "thickness2") &&
                rd.real(&out->R1, "R1") && rd.real(&out->R2, "R2") && rd.real(&out->R3, "R3") &&
                rd.real(&out->diameter, "diameter") && rd.real(&out->f, "f") && rd.real(&out->fb, "fb") &&
                rd.real(&out->n1, "n1");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&B1[i], "b(glass 1)");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&C1[i], "c(glass 1)");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&B2[i], "b(glass 2)");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&C2[i], "c(glass 2)");
    if (!good) return ORT_EPARSE;
    out->n2 = sellmeier(wavelength, B1, C1);
    out->n3 = sellmeier(wavelength, B2, C2);
    out->radius = out->diameter / 2.0;
    out->thickness = out->thickness1 + out->thickness2;
    out->centre1[2] = offset + out->fb + out->R1;
    out->centre2[2] = offset + out->fb + out->thickness1 - out->R2;
    out->centre3[2] = offset + out->fb + out->thickness - out->R3;
    return ORT_OK;
}

/* init_bottle, src/lens.f90:170-227.  The reference aborts on a 13..15-line file (an unguarded
 * read hits EOF); here missing mu lines mean 0. */
int ort_load_bottle(const char* path, double wavelength, ort_bottle* out) {
    if (!path || !out) return ORT_EINVAL;
    ListReader rd(path);
    if (!rd.ok()) {
        ort_set_error("cannot open %s", path);
        return ORT_EIO;
    }
    double g[3], a[3];
    std::memset(out, 0, sizeof *out);
    bool good = rd.real(&out->thickness, "thickness") && rd.real(&out->radiusa, "radiusa") &&
                rd.real(&out->radiusb, "radiusb") && rd.real(&out->centre[0], "x") &&
                rd.real(&out->centre[1], "y") && rd.real(&out->centre[2], "z");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&g[i], "glass coefficient");
    for (int i = 0; good && i < 3; ++i) good = rd.real(&a[i], "contents coefficient");
    if (!good) return ORT_EPARSE;
    double* mu[4] = {&out->mua_b, &out->mus_b, &out->mua_c, &out->mus_c};
    for (int i = 0; i < 4 && rd.remaining() > 0; ++i)
        if (!rd.real(mu[i], "mu")) return ORT_EPARSE;
    out->nbottle = glass_dispersion(wavelength, g[0], g[1], g[2]);
    out->ncontents = cauchy_index(wavelength, a[0], a[1], a[2]);
    out->scatter_b = (out->mua_b + out->mus_b != 0.0) ? 1 : 0;
    out->scatter_c = (out->mua_c + out->mus_c != 0.0) ? 1 : 0;
    out->ellipse = (out->radiusa != out->radiusb) ? 1 : 0;
    return ORT_OK;
}

/* read_settings, src/setupMod.f90:56-133 */
int ort_read_settings(const char* path, ort_settings* out) {
    if (!path || !out) return ORT_EINVAL;
    ListReader rd(path);
    if (!rd.ok()) {
        ort_set_error("cannot open %s", path);
        return ORT_EIO;
    }
    std::memset(out, 0, sizeof *out);
    bool good = rd.real(&out->ring_width, "ringWidth") && rd.real(&out->wavelength, "wavelength") &&
                rd.integer(&out->nphotons, "nphotons") && rd.real(&out->alpha_deg, "alpha") &&
                rd.real(&out->n_axicon, "n") && rd.logical(&out->use_bottle, "use_bottle") &&
                rd.logical(&out->use_tracker, "use_tracker") && rd.logical(&out->make_images, "makeImages") &&
                rd.real(&out->image_diameter, "image_diameter") && rd.real(&out->fibre_offset, "fibre_offset") &&
                rd.text(out->source_type, sizeof out->source_type, "source_type") &&
                rd.text(out->iris_name, sizeof out->iris_name, "iris") &&
                rd.real(&out->iris_radius, "iris_radius") &&
                rd.text(out->bottle_file, sizeof out->bottle_file, "bottle file") &&
                rd.text(out->l2_file, sizeof out->l2_file, "L2 file") &&
                rd.text(out->l3_file, sizeof out->l3_file, "L3 file") &&
                rd.text(out->image_file, sizeof out->image_file, "image source file") &&
                rd.text(out->folder, sizeof out->folder, "data folder") &&
                rd.real(&out->isors_offset, "isors_offset") && rd.real(&out->spot_size, "spot_size");
    if (!good) return ORT_EPARSE;
    const char* known[] = {"image", "spot", "point", "isors", "crs"};
    bool src_ok = false;
    for (const char* k : known) src_ok |= (std::strcmp(out->source_type, k) == 0);
    if (!src_ok) { /* src/setupMod.f90:98 */
        ort_set_error("No such source type! (%s)", out->source_type);
        return ORT_EPARSE;
    }
    if (std::strcmp(out->iris_name, "before") == 0) {
        out->iris_before = 1;
    } else if (std::strcmp(out->iris_name, "after") == 0) {
        out->iris_after = 1;
    } else if (std::strcmp(out->iris_name, "none") != 0) { /* src/setupMod.f90:110 */
        ort_set_error("No such iris position! (%s)", out->iris_name);
        return ORT_EPARSE;
    }
    if (out->nphotons < 0) {
        ort_set_error("negative number of photons");
        return ORT_EPARSE;
    }
    return ORT_OK;
}

int ort_build_scene(const ort_settings* st, const char* resdir, double lens_wavelength, ort_scene* out,
                    double* pre_guard_offset) {
    if (!st || !out) return ORT_EINVAL;
    std::memset(out, 0, sizeof *out);
    int rc = ort_load_bottle(join(resdir, st->bottle_file).c_str(), st->wavelength, &out->bottle);
    if (rc) return rc;
    rc = ort_load_plano(join(resdir, st->l2_file).c_str(), lens_wavelength, 0.0, &out->L2);
    if (rc) return rc;
    /* src/setupMod.f90:119 and src/main.f90:116: L3 sits 2*fb + thickness of L2 downstream */
    rc = ort_load_doublet(join(resdir, st->l3_file).c_str(), lens_wavelength,
                          2. * out->L2.fb + out->L2.thickness, &out->L3);
    if (rc) return rc;
    if (pre_guard_offset) *pre_guard_offset = out->bottle.centre[2];
    { /* src/setupMod.f90:135-136: the crs spot radius is rescaled before main's offset guard */
        double offset = out->bottle.radiusa + out->bottle.centre[2];
        out->spot_size = (st->spot_size * (out->L2.fb - offset)) / out->L2.fb;
        out->isors_offset = st->isors_offset;
        out->ring_width = st->ring_width;
    }

    const double pi = 3.14159265358979323846;
    const bool isors = std::strcmp(st->source_type, "isors") == 0;
    double alpha = st->alpha_deg * pi / 180.; /* src/setupMod.f90:61 */
    out->cos_theta_max = std::cos(std::atan(out->L2.radius / out->L2.fb)); /* src/main.f90:51-52 */
    if (out->L2.fb <= out->bottle.radiusa + out->bottle.centre[2]) {       /* src/main.f90:54-58 */
        out->bottle.centre[2] = out->L2.fb - out->bottle.radiusa - 2e-3;
    }
    double distance = isors ? out->bottle.radiusa + st->isors_offset
                            : (out->bottle.radiusa + out->bottle.centre[2]); /* :60-64 */
    double bessel = distance * 97.3e-3 * std::tan(alpha * (st->n_axicon - 1)) / (out->L2.fb); /* :66 */
    double inner = bessel - st->ring_width;
    out->r2 = (bessel / 2.0) * (bessel / 2.0);
    out->r1 = inner * inner;
    out->img_plane = 2. * (out->L2.fb + out->L3.fb) + out->L2.thickness + out->L3.thickness; /* :81 */
    out->point_offset = isors ? out->bottle.centre[2] : 0.0;                                  /* :140 */
    return ORT_OK;
}

int ort_job_from_settings(const ort_settings* st, int32_t phase, ort_job* out) {
    if (!st || !out) return ORT_EINVAL;
    std::memset(out, 0, sizeof *out);
    out->phase = phase;
    out->use_bottle = st->use_bottle;
    out->iris_before = st->iris_before;
    out->iris_after = st->iris_after;
    out->precision = 64;
    out->iris_radius = st->iris_radius;
    out->fibre_offset = st->fibre_offset;
    out->image_diameter = st->image_diameter;
    out->uniform_override = -1.0;
    if (std::strcmp(st->source_type, "crs") == 0) out->source_kind = ORT_SRC_CRS;
    else if (std::strcmp(st->source_type, "isors") == 0) out->source_kind = ORT_SRC_ISORS;
    else if (std::strcmp(st->source_type, "spot") == 0) out->source_kind = ORT_SRC_SPOT;
    else if (std::strcmp(st->source_type, "image") == 0) out->source_kind = ORT_SRC_IMAGE;
    else out->source_kind = ORT_SRC_POINT;
    out->total_rays = st->nphotons;
    out->seed = 123456789ull; /* src/main.f90:79 */
    out->first_ray = 0;
    out->nrays = st->nphotons;
    return ORT_OK;
}

/* init_emit_image, src/sourceMod.f90:363-408 */
int ort_load_image_source(const char* path, int64_t nphotons, uint64_t seed, int32_t* budget) {
    if (!path || !budget || nphotons < 0) return ORT_EINVAL;
    const int N = ORT_SRCIMG_N;
    std::vector<double> f((size_t)N * N);
    FILE* fh = std::fopen(path, "rb");
    if (!fh) {
        ort_set_error("cannot open image source %s", path);
        return ORT_EIO;
    }
    size_t got = std::fread(f.data(), sizeof(double), f.size(), fh);
    std::fclose(fh);
    if (got != f.size()) {
        ort_set_error("%s: expected %d x %d float64 values", path, N, N);
        return ORT_EPARSE;
    }
    /* the file is read column-major and transposed: imgout(i,j) = f[(i-1)*N + (j-1)];
     * sum() walks the transposed array in its memory order (j outer, i inner) */
    double tot = 0.0;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) tot += f[(size_t)i * N + j];
    OrtRng g;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.rk = nullptr;
    g.r1 = 0; g.phase = 3; g.override_u = -1.0;
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) {
            double share = ((double)nphotons * f[(size_t)i * N + j]) / tot;
            long long whole = (long long)share;
            double frac = share - (double)whole;
            if (whole + 1 > (long long)INT32_MAX) {
                /* the reference cannot get here (its nphotons is int32); a share this large would
                 * wrap and be dropped silently */
                ort_set_error("ort_load_image_source: pixel (%d,%d) would get %lld rays, more than an int32 budget "
                              "holds; split the job", i + 1, j + 1, whole + 1);
                return ORT_EINVAL;
            }
            g.r0 = (uint32_t)(i * N + j);
            const double u = ort_slot<double>(g, 0u);
            budget[(size_t)j * N + i] = (int32_t)((u < frac && frac > 0) ? whole + 1 : whole);
        }
    }
    return ORT_OK;
}

/* src/main.f90:45-48 */
int ort_output_basename(const ort_settings* st, const ort_scene* sc, double pre_guard_offset, char* buf,
                        size_t buflen) {
    if (!st || !sc || !buf) return ORT_EINVAL;
    const double pi = 3.14159265358979323846;
    double alpha = st->alpha_deg * pi / 180.;
    std::ostringstream o;
    o << st->source_type << "_bottle_" << (st->use_bottle ? "T" : "F") << "_Ra_" << str_real(sc->bottle.radiusa, 7)
      << "_Rb_" << str_real(sc->bottle.radiusb, 7) << "_offset_" << str_real(pre_guard_offset, 7) << "_"
      << "_" << (st->iris_before ? "T" : "F") << "_" << (st->iris_after ? "T" : "F") << "_"
      << str_real(st->iris_radius, 7) << "_L2f_" << str_real(sc->L2.f, 6) << "_L3f_" << str_real(sc->L3.f, 6)
      << "_fo_" << str_real(st->fibre_offset, 7) << "_alp_" << str_real(alpha * 180 / pi, 7) << "_bwidth_"
      << str_real(st->ring_width, 7) << "_sep_" << str_real(st->isors_offset, 7);
    std::string s = o.str();
    if (s.size() + 1 > buflen) {
        ort_set_error("output name does not fit");
        return ORT_EINVAL;
    }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return ORT_OK;
}

/* writeImage2D, src/imageMod.f90:93-114: raw fp64, x fastest */
int ort_write_images(const char* base, const uint64_t* ring, const uint64_t* point) {
    if (!base || !ring || !point) return ORT_EINVAL;
    std::vector<double> tmp(ORT_IMG_BINS);
    const char* suffix[3] = {"-ring.dat", "-point.dat", "-total.dat"};
    for (int k = 0; k < 3; ++k) {
        for (int i = 0; i < ORT_IMG_BINS; ++i) {
            double a = (double)ring[i], b = (double)point[i];
            tmp[i] = (k == 0) ? a : (k == 1) ? b : a + b;
        }
        std::string name = std::string(base) + suffix[k];
        FILE* fh = std::fopen(name.c_str(), "wb");
        if (!fh) {
            ort_set_error("cannot write %s", name.c_str());
            return ORT_EIO;
        }
        size_t w = std::fwrite(tmp.data(), sizeof(double), tmp.size(), fh);
        std::fclose(fh);
        if (w != tmp.size()) {
            ort_set_error("short write on %s", name.c_str());
            return ORT_EIO;
        }
    }
    return ORT_OK;
}

/* writeImage3D, src/imageMod.f90:117-133: real(image(:,:,:,layer)) as one raw stream per layer */
int ort_write_volume(const char* base, const uint32_t* vol_ring, const uint32_t* vol_point) {
    if (!base) return ORT_EINVAL;
    const uint32_t* vols[2] = {vol_ring, vol_point};
    const char* suffix[2] = {"-vol-ring.dat", "-vol-point.dat"};
    std::vector<double> tmp(ORT_IMG_BINS);
    for (int k = 0; k < 2; ++k) {
        if (!vols[k]) continue;
        std::string name = std::string(base) + suffix[k];
        FILE* fh = std::fopen(name.c_str(), "wb");
        if (!fh) {
            ort_set_error("cannot write %s", name.c_str());
            return ORT_EIO;
        }
        bool ok = true;
        for (int z = 0; z < ORT_VOL_DEPTH && ok; ++z) {
            const uint32_t* slab = vols[k] + (size_t)z * ORT_IMG_BINS;
            for (int i = 0; i < ORT_IMG_BINS; ++i) tmp[i] = (double)slab[i];
            ok = std::fwrite(tmp.data(), sizeof(double), tmp.size(), fh) == tmp.size();
        }
        if (std::fclose(fh) != 0) ok = false;
        if (!ok) {
            ort_set_error("short write on %s", name.c_str());
            return ORT_EIO;
        }
    }
    return ORT_OK;
}

/* src/main.f90:168-178 */
int ort_append_trans_stats(const char* folder, const ort_settings* st, const ort_scene* sc, int64_t rcount,
                           int64_t pcount) {
    if (!folder || !st || !sc) return ORT_EINVAL;
    std::string name = join(folder, "trans-stats.dat");
    struct stat sb;
    bool exists = ::stat(name.c_str(), &sb) == 0;
    FILE* fh = std::fopen(name.c_str(), exists ? "a" : "w");
    if (!fh) {
        ort_set_error("cannot write %s", name.c_str());
        return ORT_EIO;
    }
    if (!exists)
        std::fprintf(fh,
                     " r/%%, p/%%, l2%%f, l3%%f, bottle?, radiusA, radiusB, iris_pos, iris_radius, offset, "
                     "source_type, seperation\n");
    double np = (double)st->nphotons;
    double rp = 100. * (1. - (rcount / np)), pp = 100. * (1. - (pcount / np));
    /* one list-directed record: items separated by the "," literals of the reference's write */
    std::ostringstream o;
    o << " " << list_real(rp) << "," << list_real(pp) << "," << list_real(sc->L2.f) << "," << list_real(sc->L3.f)
      << ", " << (st->use_bottle ? "T" : "F") << "," << list_real(sc->bottle.radiusa) << ","
      << list_real(sc->bottle.radiusb) << ", " << (st->iris_before ? "T" : "F") << " "
      << (st->iris_after ? "T" : "F") << "," << str_real(st->iris_radius, 7) << ","
      << list_real(sc->bottle.centre[2]) << "," << st->source_type << "," << list_real(st->isors_offset);
    std::fprintf(fh, "%s\n", o.str().c_str());
    std::fclose(fh);
    return ORT_OK;
}

}  // extern "C"
