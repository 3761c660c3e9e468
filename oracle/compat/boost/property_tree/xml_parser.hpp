#pragma once
#include "ptree.hpp"
