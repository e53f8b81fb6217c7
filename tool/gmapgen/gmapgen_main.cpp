// gmapgen_main.cpp -- generates the grid-remapping table files of a coupled run, from the reference's own
// namelist file, through the C ABI only (include/dccm_b200.h + libdccm_b200.so; host code, no GPU needed).
//
// Stands in for the reference's program gmapgen_main (ref tool/gmapgen/gmapgen_main.f90:9-149): same namelist
// groups and keys, same defaults, same six files in the same on-disk format (one list-directed line
// `iD jD iS jS coef` per entry), so the component glue's set_mappingTable_interpCoef reads them unchanged.
//
//   gmapgen_main [--N=FILE | --namelist=FILE] [--lon-mode=0|1] [--format=text|bin] [--quiet]
//
//   --N / --namelist   namelist file (ref common/optionparser_mod.f90:86; default gmapgen.conf, ref :30)
//   --lon-mode=1       extension: generalised longitude overlap / unwrapped east neighbour for atmosphere and ocean
//                      grids with different longitudes (the reference generator stops on those)
//   --format=bin       extension: compact binary tables (dccm_table_write_bin) instead of text
//
// &PARAM_DCCM_GRID  IMA JMA KMA NMA IMO JMO KMO NMO                                   (ref :172-174, defaults :204-212)
// &PARAM_GMAPGEN    gmapfile_{AO,OA,AS,SA,OS,SO}_NAME interp_order_{AO,OA,AS,SA,OS,SO} ConservativeFlag
//                                                                                     (ref :176-190, defaults :214-224)
// A table whose file name is not given is skipped (the reference leaves such names uninitialised, ref :51-62).
// Grids: Gaussian latitudes and equally spaced longitudes for both components (the SPML axes of ref :256-307; the
// truncation wavenumbers NMA / NMO do not enter the axes), exchange grid by generate_surface_exchage_grid (:336-405).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/dccm_b200.h"

namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::fprintf(stderr, "gmapgen_main: %s\n", msg.c_str());
    std::exit(1);
}

void check(int rc, const std::string &where)
{
    if (rc != 0) die(where + ": " + dccm_last_error());
}

std::string lower(std::string s)
{
    for (char &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// ---- Fortran namelist input, as far as the two groups need it: `&group key = value, key = value ... /`,
// keys and group names case-insensitive, `!` comments, strings in single or double quotes, integers, logicals.
using Group = std::map<std::string, std::string>;

std::map<std::string, Group> read_namelist(const std::string &file)
{
    std::ifstream in(file);
    if (!in) die("cannot open namelist '" + file + "'");
    std::string text, line;
    while (std::getline(in, line)) {                    // drop comments outside quotes
        char q = 0;
        for (size_t k = 0; k < line.size(); k++) {
            char c = line[k];
            if (q) { if (c == q) q = 0; }
            else if (c == '\'' || c == '"') q = c;
            else if (c == '!') { line.resize(k); break; }
        }
        text += line + "\n";
    }
    std::map<std::string, Group> groups;
    size_t p = 0;
    while ((p = text.find('&', p)) != std::string::npos) {
        size_t e = ++p;
        while (e < text.size() && (std::isalnum((unsigned char)text[e]) || text[e] == '_')) e++;
        Group &g = groups[lower(text.substr(p, e - p))];
        p = e;
        while (p < text.size()) {
            while (p < text.size() && (std::isspace((unsigned char)text[p]) || text[p] == ',')) p++;
            if (p >= text.size() || text[p] == '/') { p++; break; }
            if (text.compare(p, 4, "&end") == 0 || text.compare(p, 4, "&END") == 0) { p += 4; break; }
            size_t k = p;
            while (k < text.size() && (std::isalnum((unsigned char)text[k]) || text[k] == '_')) k++;
            if (k == p) die("namelist '" + file + "': unexpected character '" + text.substr(p, 1) + "'");
            std::string key = lower(text.substr(p, k - p));
            p = k;
            while (p < text.size() && std::isspace((unsigned char)text[p])) p++;
            if (p >= text.size() || text[p] != '=') die("namelist '" + file + "': '=' expected after " + key);
            p++;
            while (p < text.size() && std::isspace((unsigned char)text[p])) p++;
            std::string val;
            if (p < text.size() && (text[p] == '"' || text[p] == '\'')) {
                char q = text[p++];
                size_t c = text.find(q, p);
                if (c == std::string::npos) die("namelist '" + file + "': unterminated string for " + key);
                val = text.substr(p, c - p);
                p = c + 1;
            } else {
                size_t c = p;
                while (c < text.size() && !std::isspace((unsigned char)text[c]) && text[c] != ',' && text[c] != '/') c++;
                val = text.substr(p, c - p);
                p = c;
            }
            g[key] = val;
        }
    }
    return groups;
}

void get(const Group &g, const char *key, int &v)
{
    auto it = g.find(lower(key));
    if (it == g.end()) return;
    char *e = nullptr;
    long x = std::strtol(it->second.c_str(), &e, 10);
    if (e == it->second.c_str() || *e) die(std::string("namelist: ") + key + " = '" + it->second + "' is not an integer");
    v = (int)x;
}

void get(const Group &g, const char *key, std::string &v)
{
    auto it = g.find(lower(key));
    if (it == g.end()) return;
    v = it->second;
    while (!v.empty() && v.back() == ' ') v.pop_back();      // trim(), as every use of the names does
}

void get(const Group &g, const char *key, bool &v)
{
    auto it = g.find(lower(key));
    if (it == g.end()) return;
    std::string s = lower(it->second);
    if (!s.empty() && s[0] == '.') s.erase(0, 1);
    if (s.empty() || (s[0] != 't' && s[0] != 'f')) die(std::string("namelist: ") + key + " is not a logical");
    v = s[0] == 't';
}

struct Grid {
    int im = 0, jm = 0;
    std::vector<double> lon, lat, lonwt, latwt;
};

Grid lonlat_grid(int im, int jm)                       // get_LonLatGrid, ref :256-307
{
    Grid g;
    g.im = im; g.jm = jm;
    g.lon.resize(im); g.lat.resize(jm); g.lonwt.resize(im); g.latwt.resize(jm);
    check(dccm_grid_gauss(im, jm, g.lon.data(), g.lat.data(), g.lonwt.data(), g.latwt.data()), "get_LonLatGrid");
    return g;
}

}  // namespace

int main(int argc, char **argv)
{
    std::string conf = "gmapgen.conf";                 // DEFAULT_GMAPGEN_CONFIGNML, ref :30
    int lon_mode = 0;
    bool binary = false, quiet = false;
    for (int a = 1; a < argc; a++) {
        std::string s = argv[a];
        auto val = [&](const char *opt) -> const char * {
            size_t n = std::strlen(opt);
            return s.compare(0, n, opt) == 0 && s.size() > n && s[n] == '=' ? s.c_str() + n + 1 : nullptr;
        };
        if (const char *v = val("--N")) conf = v;
        else if (const char *v = val("--namelist")) conf = v;
        else if (const char *v = val("--lon-mode")) lon_mode = std::atoi(v);
        else if (const char *v = val("--format")) binary = std::string(v) == "bin";
        else if (s == "--quiet") quiet = true;
        else if (s == "--help" || s == "-h") {
            std::printf("usage: gmapgen_main [--N=FILE|--namelist=FILE] [--lon-mode=0|1] [--format=text|bin] [--quiet]\n");
            return 0;
        } else die("unknown option '" + s + "' (see --help)");
    }

    // ---- read_config, ref :158-254
    int IMA = 64, JMA = 32, KMA = 26, NMA = 21, IMO = 64, JMO = 32, KMO = 26, NMO = 21;
    bool conservative = false;
    struct Pair { const char *tag; std::string file; int order; };
    Pair AO{"AO", "", 2}, OA{"OA", "", 2}, AS{"AS", "", 2}, SA{"SA", "", 1}, OS{"OS", "", 1}, SO{"SO", "", 1};
    if (!conf.empty()) {
        auto nml = read_namelist(conf);
        const Group &gg = nml["param_dccm_grid"], &gm = nml["param_gmapgen"];
        get(gg, "IMA", IMA); get(gg, "JMA", JMA); get(gg, "KMA", KMA); get(gg, "NMA", NMA);
        get(gg, "IMO", IMO); get(gg, "JMO", JMO); get(gg, "KMO", KMO); get(gg, "NMO", NMO);
        for (Pair *p : {&AO, &OA, &AS, &SA, &OS, &SO}) {
            get(gm, (std::string("gmapfile_") + p->tag + "_NAME").c_str(), p->file);
            get(gm, (std::string("interp_order_") + p->tag).c_str(), p->order);
        }
        get(gm, "ConservativeFlag", conservative);
    }
    if (!quiet) {
        std::printf(" *** MESSAGE [gmapgen_main] ***  ATM: (IM, JM, KM)=(%d,%d,%d)\n", IMA, JMA, KMA);
        std::printf(" *** MESSAGE [gmapgen_main] ***  OCN: (IM, JM, KM)=(%d,%d,%d)\n", IMO, JMO, KMO);
        std::printf(" *** MESSAGE [gmapgen_main] ***  conservativeFlag =%s\n", conservative ? "T" : "F");
        std::printf(" *** MESSAGE [gmapgen_main] ***  interp_order: (AO,OA,AS,SA,OS,SO)=(%d,%d,%d,%d,%d,%d)\n",
                    AO.order, OA.order, AS.order, SA.order, OS.order, SO.order);
    }

    // ---- grids, ref :79-87
    Grid A = lonlat_grid(IMA, JMA), O = lonlat_grid(IMO, JMO), S;
    S.im = A.im; S.lon = A.lon; S.lonwt = A.lonwt;
    S.lat.resize(JMA + JMO); S.latwt.resize(JMA + JMO);
    check(dccm_grid_exchange(A.jm, A.lat.data(), A.latwt.data(), O.jm, O.latwt.data(), &S.jm, S.lat.data(), S.latwt.data()),
          "generate_surface_exchage_grid");
    S.lat.resize(S.jm); S.latwt.resize(S.jm);
    if (!quiet) std::printf(" --Surface exchange grid-- %d x %d\n", S.im, S.jm);

    // ---- the six tables, in the reference's order AO OA AS SA OS SO (ref :103-148)
    struct Job { Pair *p; const Grid *s, *d; };
    const Job jobs[] = {{&AO, &A, &O}, {&OA, &O, &A}, {&AS, &A, &S}, {&SA, &S, &A}, {&OS, &O, &S}, {&SO, &S, &O}};
    for (const Job &j : jobs) {
        if (j.p->file.empty()) {
            if (!quiet) std::printf(" gmapfile_%s_NAME is not set: table skipped\n", j.p->tag);
            continue;
        }
        const Grid &s = *j.s, &d = *j.d;
        dccm_table *t = nullptr;
        if (conservative)
            check(dccm_table_gen_jones99(s.im, s.lon.data(), s.jm, s.lat.data(), d.im, d.lon.data(), d.jm, d.lat.data(),
                                         s.latwt.data(), d.latwt.data(), j.p->order, lon_mode, &t),
                  std::string("gen_gridmapfile_lonlat2lonlat (jones99) ") + j.p->tag);
        else
            check(dccm_table_gen_bilinear(s.im, s.lon.data(), s.jm, s.lat.data(), d.im, d.lon.data(), d.jm, d.lat.data(),
                                          lon_mode, &t),
                  std::string("gen_gridmapfile_lonlat2lonlat ") + j.p->tag);
        check((binary ? dccm_table_write_bin : dccm_table_write_text)(t, j.p->file.c_str()), "write " + j.p->file);
        if (!quiet) std::printf(" %s: %lld entries -> %s\n", j.p->tag, (long long)dccm_table_size(t), j.p->file.c_str());
        dccm_table_free(t);
    }

    // ---- check_mappingTable for A->O and O->A, ref :151-152, :309-334 (reads the file back, first four operations)
    const struct { Pair *p; int gnxs, gnxr; } checks[] = {{&AO, IMA, IMO}, {&OA, IMO, IMA}};
    for (const auto &c : checks) {
        if (c.p->file.empty() || quiet) continue;
        dccm_table *t = nullptr;
        check((binary ? dccm_table_read_bin : dccm_table_read_text)(c.p->file.c_str(), &t), "read " + c.p->file);
        const long long n = dccm_table_size(t);
        std::vector<int32_t> send(n), recv(n);
        std::vector<double> coef(n);
        check(dccm_table_index(t, c.gnxs, c.gnxr, send.data(), recv.data(), coef.data()), "set_mappingTable_interpCoef");
        std::printf(" * Set mapping table and coeffecient for interpolation.. file=%s\n", c.p->file.c_str());
        std::printf("   recv_index=");
        for (long long k = 0; k < 4 && k < n; k++) std::printf(" %d", recv[k]);
        std::printf("  send_index=");
        for (long long k = 0; k < 4 && k < n; k++) std::printf(" %d", send[k]);
        std::printf("\n   coef=");
        for (long long k = 0; k < 4 && k < n; k++) std::printf(" %.16g", coef[k]);
        std::printf("\n");
        dccm_table_free(t);
    }
    return 0;
}
