// oracle/shim/boost/algorithm/string.hpp -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the handful of Boost string algorithms the reference's preloop uses (Parameters.cpp:185-188, XMPI.cpp:52-90,
// ExodusModel.cpp:33,299, Source.cpp:109, ...), so that those sources compile unmodified into oracle/_ref (Makefile.main).
// Boost is not in this image.  Semantics follow the Boost documentation for the calls listed: ASCII case folding,
// trim = strip std::isspace characters, split on a character set with optional compression of adjacent separators.
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>

namespace boost {

namespace algorithm {
enum token_compress_mode_type { token_compress_on, token_compress_off };
}
using algorithm::token_compress_on;
using algorithm::token_compress_off;

struct ax_is_any_of {
    std::string set;
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline ax_is_any_of is_any_of(const std::string &s) { return ax_is_any_of{s}; }

inline bool iequals(const std::string &a, const std::string &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
inline std::string ax_lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
inline bool icontains(const std::string &a, const std::string &b) { return ax_lower(a).find(ax_lower(b)) != std::string::npos; }
template <class S> inline void to_upper(S &s) {
    for (auto &c : s) c = (char)std::toupper((unsigned char)c);
}
template <class S> inline void to_lower(S &s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
}
template <class Pred> inline void trim_if(std::string &s, Pred p) {
    size_t b = 0, e = s.size();
    while (b < e && p(s[b])) ++b;
    while (e > b && p(s[e - 1])) --e;
    s = s.substr(b, e - b);
}
inline void trim(std::string &s) {
    trim_if(s, [](char c) { return std::isspace((unsigned char)c) != 0; });
}
inline std::string trim_copy(const std::string &s) {
    std::string t(s);
    trim(t);
    return t;
}
inline void replace_first(std::string &s, const std::string &what, const std::string &with) {
    size_t p = s.find(what);
    if (p != std::string::npos) s.replace(p, what.size(), with);
}
// boost::split: tokens between separators; with token_compress_on adjacent separators count as one (leading / trailing
// separators still yield an empty first / last token, as in Boost)
template <class Pred>
inline std::vector<std::string> &split(std::vector<std::string> &out, const std::string &in, Pred p,
                                       algorithm::token_compress_mode_type mode = token_compress_off) {
    out.clear();
    std::string cur;
    size_t i = 0;
    const size_t n = in.size();
    while (true) {
        cur.clear();
        while (i < n && !p(in[i])) cur += in[i++];
        out.push_back(cur);
        if (i >= n) break;
        ++i;                                                   // the separator
        if (mode == token_compress_on)
            while (i < n && p(in[i])) ++i;
        if (i >= n) { out.push_back(std::string()); break; }
    }
    return out;
}

}  // namespace boost
