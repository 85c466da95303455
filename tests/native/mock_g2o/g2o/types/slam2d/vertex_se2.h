// TEST MOCK: see mock_core.h
#pragma once
#include "../../core/mock_core.h"
