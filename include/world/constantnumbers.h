// world/constantnumbers.h -- the public constants callers of the reference use next to the API
// (examples/parameter_io, examples/codec_test include this header).  Same names, same values, in
// namespace world, C++ only like the reference's header (src/world/constantnumbers.h:11-52).  The library
// itself keeps its own copies next to the code that uses them.
#ifndef WORLD_CONSTANT_NUMBERS_H_
#define WORLD_CONSTANT_NUMBERS_H_

// One row per constant: (type, name, value, what it is for).
#define WORLD_B200_PUBLIC_CONSTANTS(ROW)                                                        \
  ROW(double, kPi, 3.1415926535897932384, "pi as the reference spells it")                      \
  ROW(double, kLog2, 0.69314718055994529, "natural log of two")                                 \
  ROW(double, kEps, 0.00000000000000022204460492503131, "double precision epsilon")             \
  ROW(double, kMySafeGuardMinimum, 0.000000000001, "guard against division by zero")            \
  ROW(double, kFloorF0, 71.0, "default f0 floor: keeps the 48 kHz CheapTrick FFT at 2048")      \
  ROW(double, kCeilF0, 800.0, "default f0 ceiling")                                             \
  ROW(double, kDefaultF0, 500.0, "stands in for f0 in unvoiced frames")                         \
  ROW(double, kFloorF0StoneMask, 40.0, "lowest f0 StoneMask refines")                           \
  ROW(double, kCutOff, 50.0, "corner of DIO's low-cut filter in Hz")                            \
  ROW(double, kMaximumValue, 100000.0, "score of a rejected DIO candidate")                     \
  ROW(int, kHanning, 1, "D4C window selector")                                                  \
  ROW(int, kBlackman, 2, "D4C window selector")                                                 \
  ROW(double, kFrequencyInterval, 3000.0, "width of one aperiodicity band in Hz")               \
  ROW(double, kUpperLimit, 15000.0, "highest aperiodicity band centre in Hz")                   \
  ROW(double, kThreshold, 0.85, "default D4C LoveTrain threshold")                              \
  ROW(double, kFloorF0D4C, 47.0, "lowest f0 D4C analyses")                                      \
  ROW(double, kSafeGuardD4C, 0.000001, "D4C power floor")                                       \
  ROW(double, kM0, 1127.01048, "mel scale slope (Stevens and Volkmann 1940)")                   \
  ROW(double, kF0, 700.0, "mel scale corner in Hz")                                             \
  ROW(double, kFloorFrequency, 40.0, "lowest frequency of the coded envelope in Hz")            \
  ROW(double, kCeilFrequency, 20000.0, "highest frequency of the coded envelope in Hz")

namespace world {
#define WORLD_B200_DEFINE_CONSTANT(type, name, value, note) const type name = value;
WORLD_B200_PUBLIC_CONSTANTS(WORLD_B200_DEFINE_CONSTANT)
#undef WORLD_B200_DEFINE_CONSTANT
}  // namespace world

#endif  // WORLD_CONSTANT_NUMBERS_H_
