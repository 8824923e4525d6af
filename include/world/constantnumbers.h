// world/constantnumbers.h -- the reference's public constants (src/world/constantnumbers.h:11-52), same names
// and values, for callers that use them next to the API (examples/parameter_io, examples/codec_test).  C++ only,
// like the reference's.  The library itself carries its own copies next to the code that uses them.
#ifndef WORLD_CONSTANT_NUMBERS_H_
#define WORLD_CONSTANT_NUMBERS_H_

namespace world {
// F0 estimators
const double kCutOff = 50.0;              // DIO low-cut corner [Hz]
const double kFloorF0StoneMask = 40.0;
const double kFloorF0 = 71.0;             // 71 Hz keeps the CheapTrick FFT at 2048 points for fs = 48 kHz
const double kCeilF0 = 800.0;
const double kDefaultF0 = 500.0;          // stands in for unvoiced frames
const double kMaximumValue = 100000.0;    // score of a rejected DIO candidate
// arithmetic
const double kPi = 3.1415926535897932384;
const double kMySafeGuardMinimum = 0.000000000001;
const double kEps = 0.00000000000000022204460492503131;
const double kLog2 = 0.69314718055994529;
// D4C
const int kHanning = 1;
const int kBlackman = 2;
const double kFrequencyInterval = 3000.0;
const double kUpperLimit = 15000.0;
const double kThreshold = 0.85;
const double kFloorF0D4C = 47.0;
const double kSafeGuardD4C = 0.000001;
// codec (mel scale of Stevens & Volkmann, 1940)
const double kM0 = 1127.01048;
const double kF0 = 700.0;
const double kFloorFrequency = 40.0;
const double kCeilFrequency = 20000.0;
}  // namespace world

#endif  // WORLD_CONSTANT_NUMBERS_H_
