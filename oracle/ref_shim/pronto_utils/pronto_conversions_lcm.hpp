#pragma once
#include <conversions/pronto_conversions_lcm.hpp>
