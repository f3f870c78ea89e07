// stand-in: bot_core_ins_t is not used by rbis.cpp
