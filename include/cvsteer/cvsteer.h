// Namespace macros of the cvsteer library surface (reference cvsteer/cvsteer.h:12-15): namespace fa = Freeman & Adelson.
#ifndef CVSTEER_B200_CVSTEER_H_
#define CVSTEER_B200_CVSTEER_H_

#define _STEER_BEGIN \
    namespace fa     \
    {
#define _STEER_END }

#endif
