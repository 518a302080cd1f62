// oracle/shim/netcdf.h -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the NetCDF C API calls of the reference's NetCDF_Reader / NetCDF_Writer (S/preloop/utilities/netcdf/), so that
// those sources -- and with them ExodusModel.cpp, NuWisdom.cpp, the recorders -- compile unmodified into oracle/_ref
// (Makefile.main).  libnetcdf / libhdf5 are not in this image.  A "NetCDF file" <name> is served from the flat container
// <name>.ncflat next to it, which oracle/nc_flatten.py writes from the real NetCDF-4 / HDF5 file with the repo's own HDF5 reader:
//     "NCFLAT1\n" <nvar>\n  then per variable:  <name> <type> <ndim> <dim0> ... <nbytes>\n <raw little-endian bytes>
// with <type> one of i1 c1 i4 i8 f4 f8.  nc_get_var copies the stored bytes without conversion, exactly what the real call
// does for a matching type (NetCDF_Reader.h:45 relies on that).  Files created through nc_create are written in the same
// container on nc_close (groups become "group/" name prefixes, attributes are kept as variables named "var@att").
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

typedef int nc_type;
#define NC_NOERR 0
#define NC_BYTE 1
#define NC_CHAR 2
#define NC_SHORT 3
#define NC_INT 4
#define NC_LONG NC_INT
#define NC_FLOAT 5
#define NC_DOUBLE 6
#define NC_INT64 10
#define NC_GLOBAL (-1)
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_NETCDF4 0x1000
#define NC_MPIIO 0x2000
#define NC_INDEPENDENT 0
#define NC_COLLECTIVE 1
#define AX_NC_EBAD (-33)

struct ax_nc_var {
    std::string name, type;
    std::vector<size_t> dims;
    std::vector<char> bytes;
};
struct ax_nc_file {
    std::string path;
    bool writable = false, dirty = false;
    std::vector<ax_nc_var> vars;
    std::vector<size_t> dimlen;                 // dimension id -> length (one id per (variable, axis) for files read)
    std::vector<std::string> dimname;
    std::vector<std::vector<int>> vardims;      // variable id -> dimension ids
    std::vector<std::string> groups;            // group id g >= 1 -> prefix; ncid = file * 64 + group
};
inline std::vector<ax_nc_file *> &ax_nc_files() {
    static std::vector<ax_nc_file *> f;
    return f;
}
inline ax_nc_file *ax_nc_get(int ncid) {
    const int f = ncid / 64;
    if (ncid < 0 || f >= (int)ax_nc_files().size()) return nullptr;
    return ax_nc_files()[f];
}
inline std::string ax_nc_prefix(ax_nc_file *f, int ncid) { return (ncid % 64) ? f->groups[ncid % 64] : std::string(); }
inline size_t ax_nc_tsize(const std::string &t) { return (size_t)(t[1] - '0'); }
inline const char *ax_nc_tname(nc_type t) {
    switch (t) {
    case NC_BYTE: return "i1";
    case NC_CHAR: return "c1";
    case NC_INT: return "i4";
    case NC_FLOAT: return "f4";
    case NC_DOUBLE: return "f8";
    case NC_INT64: return "i8";
    }
    return "";
}
inline int ax_nc_register(ax_nc_file *f) {
    f->groups.assign(1, std::string());
    ax_nc_files().push_back(f);
    return ((int)ax_nc_files().size() - 1) * 64;
}
inline bool ax_nc_load(ax_nc_file *f) {
    FILE *fp = std::fopen((f->path + ".ncflat").c_str(), "rb");
    if (!fp) return false;
    char line[4096];
    if (!std::fgets(line, sizeof line, fp) || std::strncmp(line, "NCFLAT1", 7)) { std::fclose(fp); return false; }
    int nvar = 0;
    if (!std::fgets(line, sizeof line, fp) || std::sscanf(line, "%d", &nvar) != 1) { std::fclose(fp); return false; }
    for (int v = 0; v < nvar; ++v) {
        if (!std::fgets(line, sizeof line, fp)) { std::fclose(fp); return false; }
        ax_nc_var var;
        char name[2048], type[16];
        int ndim = 0, pos = 0;
        if (std::sscanf(line, "%2047s %15s %d%n", name, type, &ndim, &pos) != 3) { std::fclose(fp); return false; }
        var.name = name;
        var.type = type;
        const char *p = line + pos;
        std::vector<int> ids;
        for (int d = 0; d < ndim; ++d) {
            unsigned long long len = 0;
            int adv = 0;
            std::sscanf(p, "%llu%n", &len, &adv);
            p += adv;
            var.dims.push_back((size_t)len);
            ids.push_back((int)f->dimlen.size());
            f->dimlen.push_back((size_t)len);
            f->dimname.push_back(std::string());
        }
        unsigned long long nbytes = 0;
        std::sscanf(p, "%llu", &nbytes);
        var.bytes.resize((size_t)nbytes);
        if (nbytes && std::fread(var.bytes.data(), 1, (size_t)nbytes, fp) != (size_t)nbytes) { std::fclose(fp); return false; }
        f->vars.push_back(var);
        f->vardims.push_back(ids);
    }
    std::fclose(fp);
    return true;
}
inline void ax_nc_store(ax_nc_file *f) {
    FILE *fp = std::fopen((f->path + ".ncflat").c_str(), "wb");
    if (!fp) return;
    std::fprintf(fp, "NCFLAT1\n%d\n", (int)f->vars.size());
    for (const ax_nc_var &v : f->vars) {
        std::fprintf(fp, "%s %s %d", v.name.c_str(), v.type.c_str(), (int)v.dims.size());
        for (size_t d : v.dims) std::fprintf(fp, " %llu", (unsigned long long)d);
        std::fprintf(fp, " %llu\n", (unsigned long long)v.bytes.size());
        if (!v.bytes.empty()) std::fwrite(v.bytes.data(), 1, v.bytes.size(), fp);
    }
    std::fclose(fp);
}

inline int nc_open(const char *path, int mode, int *ncid) {
    ax_nc_file *f = new ax_nc_file;
    f->path = path;
    f->writable = (mode & NC_WRITE) != 0;
    if (!ax_nc_load(f)) { delete f; return AX_NC_EBAD; }
    *ncid = ax_nc_register(f);
    return NC_NOERR;
}
inline int nc_create(const char *path, int, int *ncid) {
    ax_nc_file *f = new ax_nc_file;
    f->path = path;
    f->writable = f->dirty = true;
    *ncid = ax_nc_register(f);
    return NC_NOERR;
}
inline int nc_sync(int ncid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    if (f->writable && f->dirty) ax_nc_store(f);
    return NC_NOERR;
}
inline int nc_close(int ncid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    if (f->writable && f->dirty) ax_nc_store(f);
    delete f;
    ax_nc_files()[ncid / 64] = nullptr;
    return NC_NOERR;
}
inline int nc_redef(int ncid) { return ax_nc_get(ncid) ? NC_NOERR : AX_NC_EBAD; }
inline int nc_enddef(int ncid) { return ax_nc_get(ncid) ? NC_NOERR : AX_NC_EBAD; }
inline int nc_inq_varid(int ncid, const char *name, int *varid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    const std::string full = ax_nc_prefix(f, ncid) + name;
    for (size_t v = 0; v < f->vars.size(); ++v)
        if (f->vars[v].name == full) { *varid = (int)v; return NC_NOERR; }
    return AX_NC_EBAD;
}
inline int nc_inq_varndims(int ncid, int varid, int *ndims) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size()) return AX_NC_EBAD;
    *ndims = (int)f->vars[varid].dims.size();
    return NC_NOERR;
}
inline int nc_inq_vardimid(int ncid, int varid, int *dimids) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size()) return AX_NC_EBAD;
    for (size_t d = 0; d < f->vardims[varid].size(); ++d) dimids[d] = f->vardims[varid][d];
    return NC_NOERR;
}
inline int nc_inq_dimlen(int ncid, int dimid, size_t *len) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || dimid < 0 || dimid >= (int)f->dimlen.size()) return AX_NC_EBAD;
    *len = f->dimlen[dimid];
    return NC_NOERR;
}
inline int nc_inq_dimid(int ncid, const char *name, int *dimid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    for (size_t d = 0; d < f->dimname.size(); ++d)
        if (!f->dimname[d].empty() && f->dimname[d] == name) { *dimid = (int)d; return NC_NOERR; }
    return AX_NC_EBAD;
}
inline int nc_def_dim(int ncid, const char *name, size_t len, int *dimid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    f->dimlen.push_back(len);
    f->dimname.push_back(name);
    *dimid = (int)f->dimlen.size() - 1;
    return NC_NOERR;
}
inline int nc_def_var(int ncid, const char *name, nc_type type, int ndims, const int *dimids, int *varid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    ax_nc_var v;
    v.name = ax_nc_prefix(f, ncid) + name;
    v.type = ax_nc_tname(type);
    if (v.type.empty()) return AX_NC_EBAD;
    size_t total = 1;
    std::vector<int> ids;
    for (int d = 0; d < ndims; ++d) {
        v.dims.push_back(f->dimlen[dimids[d]]);
        ids.push_back(dimids[d]);
        total *= f->dimlen[dimids[d]];
    }
    v.bytes.assign(total * ax_nc_tsize(v.type), 0);
    f->vars.push_back(v);
    f->vardims.push_back(ids);
    f->dirty = true;
    *varid = (int)f->vars.size() - 1;
    return NC_NOERR;
}
inline int nc_get_var(int ncid, int varid, void *out) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size()) return AX_NC_EBAD;
    if (!f->vars[varid].bytes.empty()) std::memcpy(out, f->vars[varid].bytes.data(), f->vars[varid].bytes.size());
    return NC_NOERR;
}
inline int nc_get_var_text(int ncid, int varid, char *out) { return nc_get_var(ncid, varid, out); }
inline int nc_put_var(int ncid, int varid, const void *in) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size()) return AX_NC_EBAD;
    if (!f->vars[varid].bytes.empty()) std::memcpy(f->vars[varid].bytes.data(), in, f->vars[varid].bytes.size());
    f->dirty = true;
    return NC_NOERR;
}
inline int nc_put_vara(int ncid, int varid, const size_t *start, const size_t *count, const void *in) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || varid < 0 || varid >= (int)f->vars.size()) return AX_NC_EBAD;
    ax_nc_var &v = f->vars[varid];
    const size_t ts = ax_nc_tsize(v.type), nd = v.dims.size();
    size_t total = 1;
    for (size_t d = 0; d < nd; ++d) {
        if (start[d] + count[d] > v.dims[d]) return AX_NC_EBAD;
        total *= count[d];
    }
    std::vector<size_t> idx(nd, 0);
    const char *src = (const char *)in;
    for (size_t n = 0; n < total; ++n) {
        size_t off = 0;
        for (size_t d = 0; d < nd; ++d) off = off * v.dims[d] + start[d] + idx[d];
        std::memcpy(v.bytes.data() + off * ts, src + n * ts, ts);
        for (size_t d = nd; d-- > 0;) {
            if (++idx[d] < count[d]) break;
            idx[d] = 0;
        }
    }
    f->dirty = true;
    return NC_NOERR;
}
inline int ax_nc_put_att(int ncid, int varid, const char *name, const char *type, size_t len, const void *val) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    ax_nc_var v;
    v.name = (varid == NC_GLOBAL ? std::string() : f->vars[varid].name) + "@" + name;
    v.type = type;
    v.dims.assign(1, len);
    v.bytes.assign((const char *)val, (const char *)val + len * ax_nc_tsize(v.type));
    f->vars.push_back(v);
    f->vardims.push_back(std::vector<int>());
    f->dirty = true;
    return NC_NOERR;
}
inline int nc_put_att(int ncid, int varid, const char *name, nc_type type, size_t len, const void *val) {
    return ax_nc_put_att(ncid, varid, name, ax_nc_tname(type), len, val);
}
inline int nc_put_att_text(int ncid, int varid, const char *name, size_t len, const char *val) {
    return ax_nc_put_att(ncid, varid, name, "c1", len, val);
}
inline int nc_def_grp(int ncid, const char *name, int *grpid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f || f->groups.size() >= 64) return AX_NC_EBAD;
    f->groups.push_back(ax_nc_prefix(f, ncid) + name + "/");
    *grpid = (ncid / 64) * 64 + (int)f->groups.size() - 1;
    return NC_NOERR;
}
inline int nc_inq_grp_ncid(int ncid, const char *name, int *grpid) {
    ax_nc_file *f = ax_nc_get(ncid);
    if (!f) return AX_NC_EBAD;
    const std::string full = ax_nc_prefix(f, ncid) + name + "/";
    for (size_t g = 1; g < f->groups.size(); ++g)
        if (f->groups[g] == full) { *grpid = (ncid / 64) * 64 + (int)g; return NC_NOERR; }
    return AX_NC_EBAD;
}
