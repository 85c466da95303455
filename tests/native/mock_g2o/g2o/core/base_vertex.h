// TEST MOCK: see mock_core.h
#pragma once
#include "mock_core.h"
