// model.cfg (de)serialiser.  Mirrors the reference's Model (Model.h:8-19): Model(string), Save, Load returning
// EXIT_SUCCESS / EXIT_FAILURE, never throwing.  The file format is libconfig 1.4.9 text with the schema of
// Model.cpp:26-79; parsing and printing are this repo's own (cfgfile.h).
#ifndef MODEL_H
#define MODEL_H

#include <string>

class CascadeClassifier;

class Model
{
public:
    std::string model_cfg;

    Model(std::string model_cfg) : model_cfg(model_cfg) {}
    ~Model();
    int Save(CascadeClassifier& cascade_classifier);
    int Load(CascadeClassifier& cascade_classifier);
};

#endif
