// wemu.h -- host lane emulator (TEST INFRASTRUCTURE), see wemu.cpp and warp_ops.cuh
#ifndef WEMU_H
#define WEMU_H
#include <functional>
unsigned long long wemu_exchange(unsigned long long v, int srcLane, int line);
unsigned wemu_ballot(int pred, int line);
int wemu_lane(void);
void wemu_run(const std::function<void(int)>& body);   // runs body(lane) on 32 lanes to completion
extern unsigned long long wemu_collectives;
#endif
