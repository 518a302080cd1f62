// oracle/shim/boost/lexical_cast.hpp -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for boost::lexical_cast as the reference's preloop uses it (Parameters.h:24-78 and a few readers): text <-> number
// through a stream, the WHOLE text must be consumed, otherwise bad_lexical_cast.  Boost is not in this image.
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace boost {

class bad_lexical_cast : public std::bad_cast {
public:
    const char *what() const noexcept override { return "bad lexical cast: source type value could not be interpreted as target"; }
};

template <class To, class From> struct ax_lexical {
    static To run(const From &x) {
        std::stringstream ss;
        ss.precision(17);
        To out;
        if (!(ss << x) || !(ss >> out)) throw bad_lexical_cast();
        char c;
        if (ss >> c) throw bad_lexical_cast();
        return out;
    }
};
template <class From> struct ax_lexical<std::string, From> {
    static std::string run(const From &x) {
        std::ostringstream ss;
        ss.precision(17);
        ss << x;
        return ss.str();
    }
};
template <> struct ax_lexical<std::string, std::string> {
    static std::string run(const std::string &x) { return x; }
};
template <class To, class From> inline To lexical_cast(const From &x) { return ax_lexical<To, From>::run(x); }

}  // namespace boost
