// installed-header name -> the reference's source tree (-I /root/reference/pronto-utils/src)
#pragma once
#include <pronto_math/pronto_math.hpp>
