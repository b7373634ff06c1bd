#pragma once
#include "IndexFlat.h"
