// test mock: see ceres/mock_all.h
#include "../ceres/mock_all.h"
