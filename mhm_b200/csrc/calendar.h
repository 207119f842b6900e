// calendar.h -- host-side calendar stepping of the library.
// Mirrors the *behaviour* of common/mo_common_datetime_type.f90:71-155 (datetimeinfo
// init / increment / update_LAI_timestep) and of the yId update in
// mHM/mo_mhm_interface_run.f90:626-628, with the date arithmetic done on a proleptic
// Gregorian day count (days-from-civil), which equals FORCES' julday/caldat for every
// date after 1582-10-15.
#pragma once
#include <cstdint>
#include <vector>

#include "device_types.h"

namespace mhm {

// Julian day number (the integer FORCES julday() returns) of a civil date
inline int64_t jdn_from_civil(int64_t y, int m, int d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468 + 2440588;
}

inline void civil_from_jdn(int64_t jdn, int& y, int& m, int& d) {
  int64_t z = jdn - 2440588 + 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const int64_t doe = z - era * 146097;
  const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t yy = yoe + era * 400;
  const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int64_t mp = (5 * doy + 2) / 153;
  d = (int)(doy - (153 * mp + 2) / 5 + 1);
  m = (int)(mp < 10 ? mp + 3 : mp - 9);
  y = (int)(yy + (m <= 2));
}

struct TimeAxis {
  int32_t jul_start = 0, nTimeSteps = 0, warming_days = 0, timeStep_LAI_input = 0;
  int32_t lc_year_start = 0;
  std::vector<int32_t> LCyearId;

  int32_t scene_of_year(int year) const {
    if (LCyearId.empty()) return 1;
    long i = (long)year - lc_year_start;
    if (i < 0) i = 0;
    if (i >= (long)LCyearId.size()) i = (long)LCyearId.size() - 1;
    return LCyearId[(size_t)i];
  }
};

// indices of steps tt = 1 .. n (the stepping is a recurrence, so always from tt = 1)
inline void fill_step_indices(const TimeAxis& ax, int timestep_h, int nTstepForcingDay, int n,
                              std::vector<StepIdx>& out) {
  out.resize((size_t)n);
  int y, m, d;
  civil_from_jdn(ax.jul_start, y, m, d);
  int hour = 0, iLAI = 0;
  bool new_day = true, new_month = true, new_year = true;
  int yId = ax.scene_of_year(y);
  // hours covered by one meteo step: nint(24 / nTstepForcingDay), mo_meteo_handler.f90:607
  const int per = (int)(24.0 / (double)nTstepForcingDay + 0.5);
  for (int tt = 1; tt <= n; ++tt) {
    switch (ax.timeStep_LAI_input) {  // update_LAI_timestep
      case 0: case 1: iLAI = m; break;
      case -1: if (new_day) ++iLAI; break;
      case -2: if (new_month) ++iLAI; break;
      case -3: if (new_year) ++iLAI; break;
      default: break;
    }
    StepIdx& s = out[(size_t)tt - 1];
    s.iMeteoTS = (tt + per - 1) / per;  // ceiling(tt / per)
    s.yId = (int16_t)yId;
    s.iLAI = (int16_t)iLAI;
    s.doy = (int16_t)(jdn_from_civil(y, m, d) - jdn_from_civil(y, 1, 1) + 1);
    s.year = (int16_t)y;
    s.month = (int8_t)m;
    s.hour = (int8_t)hour;
    s.isday = (int8_t)((hour > 6) && (hour <= 18));
    s.flags = 0;
    // increment
    const int pd = d, pm = m, py = y;
    hour += timestep_h;
    const int64_t jul = jdn_from_civil(y, m, d) + hour / 24;
    hour %= 24;
    civil_from_jdn(jul, y, m, d);
    new_day = pd != d;
    new_month = pm != m;
    new_year = py != y;
    s.flags = (int8_t)((new_day ? 1 : 0) | (new_month ? 2 : 0) | (new_year ? 4 : 0));
    if (new_year && tt < ax.nTimeSteps) yId = ax.scene_of_year(y);
  }
}

}  // namespace mhm
