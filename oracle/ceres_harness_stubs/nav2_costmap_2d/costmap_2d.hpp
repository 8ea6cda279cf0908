// TEST INFRASTRUCTURE — stand-in for nav2_costmap_2d::Costmap2D (the accessors the reference reads:
// src/optimizer.cpp:167-168, src/critics/obstacle_cost_function.cpp:24-26).
#pragma once
namespace nav2_costmap_2d {
class Costmap2D {
public:
  Costmap2D(unsigned char* data, unsigned int sx, unsigned int sy, double res, double ox, double oy)
    : data_(data), sx_(sx), sy_(sy), res_(res), ox_(ox), oy_(oy) {}
  unsigned char* getCharMap() const { return data_; }
  unsigned int getSizeInCellsX() const { return sx_; }
  unsigned int getSizeInCellsY() const { return sy_; }
  double getResolution() const { return res_; }
  double getOriginX() const { return ox_; }
  double getOriginY() const { return oy_; }
private:
  unsigned char* data_; unsigned int sx_, sy_; double res_, ox_, oy_;
};
}
