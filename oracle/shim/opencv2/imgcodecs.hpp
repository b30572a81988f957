#pragma once
#include "../ref_cv_stub.hpp"
