// SHIM (test infrastructure, ours): included by the reference's bvh.cpp:41, nothing of it is used there.
#pragma once
namespace drake { namespace geometry {} }
