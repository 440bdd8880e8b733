/* empty: blockIdx & co. come from hostemu.h — TEST INFRASTRUCTURE ONLY */
