// Headless stand-in for the reference's OpenGL viewer translation unit
// (simulation/redmax/SimViewer.cpp is not compiled: no GL/X11 here).
// Defines the members declared in simulation/redmax/SimViewer.h:11-71 as
// no-ops so that Simulation::replay() links and does nothing.
#include "SimViewer.h"

namespace redmax {

const float SimViewerTimer::CurrentTime() { return 0.f; }
void SimViewerTimer::reset() {}
void SimViewerTimer::reverse_run_or_pause() {}
void SimViewerTimer::slow_down() {}
void SimViewerTimer::speed_up() {}

void SimViewerKeyboardHandler::KeyCallback(const int, const int) {}

SimViewer::SimViewer(Simulation* sim) : _sim(sim), _timer(nullptr), _keyboard_handler(nullptr) {}
SimViewer::~SimViewer() {}
void SimViewer::initialize() {}
void SimViewer::run() {}
void SimViewer::reverse_run_or_pause() {}
void SimViewer::slow_down() {}
void SimViewer::speed_up() {}

}  // namespace redmax
