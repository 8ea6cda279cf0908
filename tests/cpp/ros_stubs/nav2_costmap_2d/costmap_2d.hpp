#pragma once
#include "ros_stub_types.hpp"
