#pragma once
#include <conversions/pronto_conversions_bot_core.hpp>
