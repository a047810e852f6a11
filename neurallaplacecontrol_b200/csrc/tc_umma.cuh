// tcgen05 / TMEM / mbarrier PTX wrappers (filled in with encode_tc.cu)
#pragma once
