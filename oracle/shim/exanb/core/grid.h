#pragma once
// stand-in: the exaNBody grid is not needed by the functors compiled here
