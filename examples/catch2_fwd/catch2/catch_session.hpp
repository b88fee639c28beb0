// forwards the modular Catch2 header name to the amalgamated distribution (thirdParty/catch2/extras of the reference tree,
// found through -I at build time; nothing of Catch2 is copied into this repository)
#pragma once
#include <catch_amalgamated.hpp>
