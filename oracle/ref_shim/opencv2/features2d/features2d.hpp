/* OpenCV stand-in (test infrastructure): see cvshim.hpp */
#include "../../cvshim.hpp"
