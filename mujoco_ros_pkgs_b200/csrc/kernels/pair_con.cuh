// pair_con.cuh — the contacts one candidate geom pair can produce, kept by the lane that owns the pair.
#pragma once

namespace b2k {

#define B2K_MAXPAIRCON 8
#define B2K_MAXPAIRFRAME 8  /* height-field pairs: every contact has its own normal */

struct PairCon {
  double dist[B2K_MAXPAIRCON];
  double pos[3 * B2K_MAXPAIRCON];
  double frame[6 * B2K_MAXPAIRFRAME];  // normal + tangent hint per contact, or one for all (shared_frame)
  int shared_frame;
};

}  // namespace b2k
