// TEST INFRASTRUCTURE — stand-in for angles/angles.h (included by goal_align_cost_function.hpp, not used by it).
#pragma once
