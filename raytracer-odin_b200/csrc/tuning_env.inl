// tuning_env.inl — included by ort_create only in -DORT_TUNING builds (make variant): the traversal / queue knobs
// that are fixed at their measured optimum in the shipped library, read from the environment for tools/tune.py.
    if (const char* e2 = std::getenv("ORT_REFILL")) c->refill = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_REFILL_LIGHT")) c->refill_light = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_INNER_MIN")) c->inner_min = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_TILED")) c->tiled = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_TILE")) {
        int a_ = 2, b_ = 2, c_ = 8;
        if (std::sscanf(e2, "%dx%dx%d", &a_, &b_, &c_) == 3 && a_ * b_ * c_ == 32) { c->tile_w = a_; c->tile_h = b_; c->tile_s = c_; }
    }
    if (const char* e2 = std::getenv("ORT_LIGHT_PREFILTER")) c->light_prefilter = std::atoi(e2);
    if (const char* e2 = std::getenv("ORT_BIN")) c->bin_octants = std::atoi(e2);
