#pragma once
#include "catch_test_macros.hpp"
