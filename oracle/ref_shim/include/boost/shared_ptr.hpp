// SHIM (test infrastructure, ours): the reference's bvh.h:54 includes <boost/shared_ptr.hpp> but uses only
// std::shared_ptr, so an empty header satisfies it.  Boost is not installed in this image.
#pragma once
#include <memory>
