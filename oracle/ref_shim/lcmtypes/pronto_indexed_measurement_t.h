// stand-in: pronto_indexed_measurement_t is not used by rbis.cpp
