// TEST STAND-IN for the reference's include/se2lam/Config.h: Config::Kcam only (Config.h:17).
#pragma once
#include "se2lam/cv_compat.h"
namespace se2lam { struct Config { static cv::Mat Kcam; }; }
