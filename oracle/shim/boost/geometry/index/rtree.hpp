// oracle/shim -- TEST INFRASTRUCTURE: see boost/geometry.hpp
#pragma once
#include <boost/geometry.hpp>
